// Host-only harness for tests/test_search_params.py: prints the launch constants of the sliding-window Search
// kernel (rtlamr_b200/csrc/search.cuh: make_slide_params) for a list of symbol lengths.  No kernel is launched.
#include <cstdio>
#include <cstdlib>

#include "../../rtlamr_b200/csrc/search.cuh"

int main(int argc, char** argv) {
    printf("[");
    bool firstrow = true;
    for (int a = 1; a + 2 < argc; a += 3) {
        const int SL = atoi(argv[a]), nbits = atoi(argv[a + 1]);
        const long long p0 = atoll(argv[a + 2]);
        ert::DevCfg c{};
        c.SL = SL;
        c.CL = SL / 2;
        c.npre = 1;
        c.pre_nbits[0] = nbits;
        for (int k = 0; k < nbits; k++) c.pre_bits[0][k] = (uint8_t)((0x1F2A60u >> (20 - (k % 21))) & 1u);
        ert::SlideParams sp;
        const bool ok = ert::make_slide_params(c, p0, 1 << 20, &sp);
        printf("%s{\"SL\":%d,\"nbits\":%d,\"p0\":%lld,\"ok\":%s,\"q\":%d,\"odd_off\":%d,\"half\":%d,\"nseg\":%d,\"tile_words\":%d,"
               "\"load_words\":%d,\"w0\":%d,\"seg\":%d,\"max_load\":%d,\"threads\":%d,\"pattern\":%u}",
               firstrow ? "" : ",", SL, nbits, p0, ok ? "true" : "false", sp.q, sp.odd_off, sp.half, sp.nseg, sp.tile_words,
               sp.load_words, sp.w0, ert::kSlideSeg, ert::kSlideMaxLoad, ert::kSlideThreads, ert::slide_pattern(c, 0));
        firstrow = false;
    }
    printf("]\n");
    return 0;
}
