// copy_pool.hpp -- host threads that copy a pageable buffer into a pinned one in parallel (the staging of pageable input in
// ertgpu_decode: a Go slice, a numpy array).  Header-only so that the CPU test (tests/test_copy_pool.py) can build it alone.
#pragma once

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace ert {

// Host threads that copy a pageable buffer into a pinned one in parallel.  Persistent: starting a std::thread costs
// 30-50 us, a 32 MiB chunk takes ~0.7 ms to copy with 8 threads -- spawning them per chunk (round 2's first form) made
// 8 threads slower than 4.  One pool per handle, created on the first pageable call.
class CopyPool {
public:
    explicit CopyPool(int nworkers) {
        for (int i = 0; i < nworkers; i++) workers_.emplace_back([this, i] { run(i + 1); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }
    // dst[0, n) = src[0, n) on the calling thread + the workers; returns when every part is done
    void copy(uint8_t* dst, const uint8_t* src, size_t n) {
        const size_t kMin = 1u << 20;
        const int parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)size(), (n + kMin - 1) / kMin));
        const size_t per = (((n + (size_t)parts - 1) / (size_t)parts) + 4095) & ~(size_t)4095;
        if (parts > 1) {
            std::lock_guard<std::mutex> g(m_);
            dst_ = dst; src_ = src; n_ = n; per_ = per; parts_ = parts;
            pending_ = parts - 1;
            gen_++;
        }
        if (parts > 1) cv_.notify_all();
        memcpy(dst, src, std::min(n, per));
        if (parts > 1) {
            std::unique_lock<std::mutex> g(m_);
            done_.wait(g, [this] { return pending_ == 0; });
        }
    }

private:
    void run(int part) {
        unsigned long long seen = 0;
        for (;;) {
            uint8_t* dst; const uint8_t* src; size_t n, per; int parts;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                dst = dst_; src = src_; n = n_; per = per_; parts = parts_;
            }
            if (part >= parts) continue;
            const size_t a = std::min(n, per * (size_t)part), b = std::min(n, per * (size_t)(part + 1));
            if (b > a) memcpy(dst + a, src + a, b - a);
            {
                std::lock_guard<std::mutex> g(m_);
                pending_--;
            }
            done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    bool stop_ = false;
    uint8_t* dst_ = nullptr;
    const uint8_t* src_ = nullptr;
    size_t n_ = 0, per_ = 0;
    int parts_ = 0, pending_ = 0;
};

}  // namespace ert
