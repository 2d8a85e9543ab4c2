// CPU harness for rtlamr_b200/csrc/copy_pool.hpp: many copies of random sizes through pools of several sizes must equal memcpy.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../rtlamr_b200/csrc/copy_pool.hpp"

int main() {
    std::vector<uint8_t> src(40u << 20), dst(40u << 20);
    uint64_t x = 88172645463325252ull;
    for (auto& b : src) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)x; }
    long long checked = 0;
    for (int workers : {0, 1, 3, 7}) {
        ert::CopyPool pool(workers);
        if (pool.size() != workers + 1) { printf("{\"ok\": false, \"why\": \"size\"}\n"); return 1; }
        for (int it = 0; it < 60; it++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            size_t n = it < 6 ? (size_t[]){0, 1, 4095, 4096, 4097, (1u << 20) + 1}[it] : (size_t)(x % src.size());
            size_t off = (size_t)((x >> 32) % (src.size() - n + 1));
            std::fill(dst.begin(), dst.end(), (uint8_t)0xA5);
            pool.copy(dst.data() + off, src.data() + off, n);
            for (size_t i = 0; i < dst.size(); i++) {
                const uint8_t want = (i >= off && i < off + n) ? src[i] : (uint8_t)0xA5;
                if (dst[i] != want) { printf("{\"ok\": false, \"workers\": %d, \"n\": %zu, \"at\": %zu}\n", workers, n, i); return 1; }
            }
            checked++;
        }
    }
    printf("{\"ok\": true, \"copies\": %lld}\n", checked);
    return 0;
}
