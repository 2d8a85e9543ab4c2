// ertgpu.cu -- C ABI of libertgpu.so (see include/ertgpu.h) and the host-side pipeline that
// drives the sm_100a kernels.  One handle mirrors one reference protocol.Decoder
// (protocol/decode.go:45-63).
//
// Device data layout (all per handle):
//   iq      : the call's interleaved uint8 I,Q bytes (caller's device buffer or a staging chunk)
//   hist[2] : last PKL samples of the previous call (ping-pong) -- replaces the Signal tail
//             (decode.go:165) and the r900 parser's private signal history (r900.go:169)
//   plane[2]: packed quantizer output, 1 bit per sample, MSB first, laid out as
//             [hist_words of history | BS/32 words per block of the call] (ping-pong) --
//             replaces Decoder.Quantized (1 byte per bit) and Decoder.packed
//   hits    : (start, preamble) pairs found by search_kernel
//   out     : ertgpu_candidate records produced by extract_kernel
#include <ctype.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "copy_pool.hpp"
#include "demod_fast.cuh"
#include "demod_generic.cuh"
#include "ert_common.cuh"
#include "search.cuh"
#include "r900_tmem.cuh"
#include "synth.cuh"

using namespace ert;

struct ertgpu_handle {
    std::string err;
    std::vector<ertgpu_protocol> protos;
    ertgpu_decoder_config cfg{};
    bool allocated = false;
    int device = 0;

    DevCfg dcfg{};
    Gf32 gf{};
    float h_lut[256];

    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};

    float* d_lut = nullptr;
    uint16_t* d_crc = nullptr;
    uint16_t* d_crc_pos = nullptr;  // per-position CRC tables of every screen
    uint32_t* d_plane[2] = {nullptr, nullptr};
    size_t plane_words = 0;
    int cur_plane = 0;
    uint8_t* d_hist[2] = {nullptr, nullptr};
    int cur_hist = 0;
    int hist_valid = 0;
    uint8_t* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    uint8_t* h_stage[2] = {nullptr, nullptr};  // pinned staging of pageable input (allocated on first use)
    size_t h_stage_bytes = 0;
    std::unique_ptr<CopyPool> copy_pool;   // host threads of the pageable-input staging
    std::string kernels;                        // instantiations launched by the last pipeline
    RawHit* d_hits = nullptr;
    HitWord* d_words = nullptr;       // words of starts that hold a hit (what Slice works from)
    uint8_t* d_digits = nullptr;
    int* d_block_slot = nullptr;      // r900: scratch slot of each block of the call (-1 none)
    int* d_slot_block = nullptr;
    unsigned int* d_slot_count = nullptr;
    float* d_r900_scratch = nullptr;  // [slot][r900_span] running sums
    int r900_slots = 0, r900_span = 0;
    ertgpu_candidate* d_out = nullptr;
    unsigned long long cand_cap = 0;
    unsigned long long* d_counters = nullptr;  // kCntHits, kCntOut, kCntValid, kCntTile (demod work tiles), kCntWords
    unsigned long long* h_counters = nullptr;  // pinned mirror
    float* d_tap = nullptr;                    // scratch for taps
    size_t tap_floats = 0;

    int64_t max_blocks = 0;
    int64_t block_counter = 0;  // number of reference Decode calls consumed so far
    bool has_r900 = false;
    int demod_variant = 0;      // 0 generic, else specialised chip length
    int demod_warps = 0;        // 0 default; tuning override (env ERTGPU_FAST_WARPS)
    bool search_legacy = false; // env ERTGPU_SEARCH_LEGACY: the per-bit-load Search kernel (kept for chip lengths like 78 and as a cross-check)
    int sm_count = 148;
    bool use_pdl = true;      // env ERTGPU_PDL=0: ordinary launches for Search and Slice
    int r900_chain_mode = 0;  // env ERTGPU_R900_CHAIN: (default) producer/consumer warps through shared memory; "tmem" = Tensor Memory between the stages; "smem" = one warp, serial lane 0; "shfl" = one warp, shuffled sums

    // state of the last enqueued pipeline (for fetch and taps)
    bool pending = false;       // a pipeline is enqueued and not yet synchronised
    bool uncopied = false;      // ... synchronised, candidates still only on the device
    unsigned long long uncopied_n = 0;
    bool fresh = true;          // next pipeline must see zeroed history (start of stream / after reset)
    cudaStream_t pending_stream = nullptr;
    const uint8_t* last_iq = nullptr;
    const uint8_t* last_hist = nullptr;
    int last_hist_valid = 0;
    const uint32_t* last_plane = nullptr;
    int64_t last_first_block = 0, last_nblocks = 0;
    bool overflow = false;
    unsigned long long need = 0;
    int64_t total_hits = 0, total_valid = 0;
    int64_t launches = 0;
    bool stage_timing = false;
    // stage timing: a pool of event sets so that nothing has to be read back between steps; the elapsed
    // times of finished pipelines are folded into the sums when the pool is full or when somebody asks
    static constexpr int kStageSets = 64;
    cudaEvent_t ev_pool[kStageSets][5] = {};
    cudaEvent_t* ev_stage = ev_pool[0];  // the set of the last enqueued pipeline
    int stage_next = 0;                  // set the next pipeline will use
    int stage_unread = 0;                // finished or pending sets not yet folded (the last stage_unread sets)
    bool stage_valid = false;
    double stage_sum[4] = {0, 0, 0, 0};  // accumulated since ertgpu_set_stage_timing(1)
    int64_t stage_n = 0;
    std::vector<ertgpu_candidate> results;
};

namespace {

int fail(ertgpu_handle* h, int code, const char* fmt, ...) {
    if (h) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}

#define CUDA_TRY(h, expr)                                                                       \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                 \
            return fail(h, e__ == cudaErrorMemoryAllocation ? ERTGPU_ENOMEM : ERTGPU_ECUDA,     \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                       \
    } while (0)

// decode.go:377-379
int32_t next_pow2(int32_t v) {
    int32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// decode.go:209-216 -- computed on the host with IEEE float32 division and multiplication
void make_maglut(float* lut) {
    for (int i = 0; i < 256; i++) {
        volatile float x = (127.5f - (float)i) / 127.5f;
        volatile float y = x * x;
        lut[i] = y;
    }
}

// crc.go:34-47
void make_crc_table(uint16_t poly, uint16_t* t) {
    for (int i = 0; i < 256; i++) {
        uint16_t crc = (uint16_t)(i << 8);
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ poly) : (uint16_t)(crc << 1);
        t[i] = crc;
    }
}

// gf.go:20-57 for NewField(32, 37, 2)
void make_gf32(Gf32* g) {
    int x = 1;
    for (int i = 0; i < 31; i++) {
        g->exp[i] = g->exp[i + 31] = (uint8_t)x;
        g->log[x] = (uint8_t)i;
        x <<= 1;
        if (x & 32) x ^= 37;
    }
    g->log[0] = 31;
}

void free_device(ertgpu_handle* h) {
    if (!h->allocated) return;
    cudaSetDevice(h->device);
    auto dfree = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    auto evfree = [](cudaEvent_t& e) { if (e) cudaEventDestroy(e); e = nullptr; };
    for (int k = 0; k < 2; k++) {
        dfree(h->d_plane[k]);
        dfree(h->d_hist[k]);
        dfree(h->d_stage[k]);
        if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]);
        h->h_stage[k] = nullptr;
        evfree(h->ev_h2d[k]);
        evfree(h->ev_done[k]);
    }
    h->stage_bytes = h->h_stage_bytes = 0;
    for (auto& set : h->ev_pool) for (int k = 0; k < 5; k++) evfree(set[k]);
    dfree(h->d_lut);
    dfree(h->d_crc);
    dfree(h->d_crc_pos);
    dfree(h->d_hits);
    dfree(h->d_words);
    dfree(h->d_digits);
    dfree(h->d_block_slot);
    dfree(h->d_slot_block);
    dfree(h->d_slot_count);
    dfree(h->d_r900_scratch);
    dfree(h->d_out);
    dfree(h->d_counters);
    dfree(h->d_tap);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    h->h_counters = nullptr;
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    h->stream = h->copy_stream = nullptr;
    h->pending = h->uncopied = false;
    h->last_plane = nullptr;
    h->allocated = false;
}

// Fold the event sets of finished pipelines into the running sums (only called with no pipeline pending).
void fold_stage_times(ertgpu_handle* h) {
    for (int i = h->stage_unread; i > 0; i--) {
        cudaEvent_t* set = h->ev_pool[(h->stage_next - i + 2 * ertgpu_handle::kStageSets) % ertgpu_handle::kStageSets];
        float ms[4];
        bool ok = true;
        for (int k = 0; k < 4; k++) ok = ok && cudaEventElapsedTime(&ms[k], set[k], set[k + 1]) == cudaSuccess;
        if (ok) {
            for (int k = 0; k < 4; k++) h->stage_sum[k] += ms[k];
            h->stage_n++;
        }
    }
    h->stage_unread = 0;
}

// Launch `kern` as a programmatic dependent of the kernel in front of it on the stream (its CTAs may become resident
// while that kernel drains; the kernel itself calls pdl_wait() before it reads anything).  pdl == false: ordinary launch.
template <class K, class... Args>
cudaError_t launch_dep(bool pdl, K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// Enqueue the whole per-call pipeline for `nblocks` blocks whose IQ bytes are at d_iq.
int enqueue_pipeline(ertgpu_handle* h, const uint8_t* d_iq, int64_t nblocks, uint32_t flags,
                     cudaStream_t st) {
    const DevCfg& c = h->dcfg;
    uint32_t* plane = h->d_plane[h->cur_plane];
    uint32_t* plane_next = h->d_plane[h->cur_plane ^ 1];
    const uint8_t* hist = h->d_hist[h->cur_hist];
    uint8_t* hist_next = h->d_hist[h->cur_hist ^ 1];
    const long long nwords = nblocks * c.words_per_block;
    const long long p0 = (long long)c.hist_words * 32 - c.PKL;
    h->launches = 0;

    CUDA_TRY(h, cudaMemsetAsync(h->d_counters, 0, kCntN * sizeof(unsigned long long), st));
    if (h->fresh) {  // Quantized starts as zeros (decode.go:145); only the history in front of the call is read
        CUDA_TRY(h, cudaMemsetAsync(plane, 0, (size_t)c.hist_words * sizeof(uint32_t), st));
        h->fresh = false;
    }

    const bool tm = h->stage_timing;
    const bool pdl = h->use_pdl && !tm;   // events recorded between the kernels would sit between primary and dependent
    if (tm) {
        if (h->stage_unread == ertgpu_handle::kStageSets) fold_stage_times(h);  // every earlier pipeline has been waited for
        h->ev_stage = h->ev_pool[h->stage_next];
        h->stage_next = (h->stage_next + 1) % ertgpu_handle::kStageSets;
        h->stage_unread++;
    }
    if (tm) CUDA_TRY(h, cudaEventRecord(h->ev_stage[0], st));
    // 1. magnitude + matched filter + quantize + pack
    char kname[160];
    if (h->demod_variant != 0) {
        int w_used = 0, var_used = 0;
        int rc = launch_demod_fast(h->demod_variant, h->demod_warps, d_iq, hist, c.hist_samples, h->hist_valid, h->d_lut,
                                   plane + c.hist_words, nblocks, c.BS, h->d_counters + kCntTile, st, &w_used, &var_used);
        if (rc != 0) return fail(h, ERTGPU_ECUDA, "demod_fast launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        snprintf(kname, sizeof(kname), "demod_fast_kernel<%d,%d,%d>", h->demod_variant, w_used, var_used);
        h->kernels = kname;
    } else {
        h->kernels = "demod_generic_kernel";
        int nthr = 128;
        while (nthr > 32 && (256 + 2 * (size_t)c.CL * nthr) * sizeof(float) > 200 * 1024) nthr >>= 1;
        const size_t smem = (256 + 2 * (size_t)c.CL * nthr) * sizeof(float);
        if (smem > 227 * 1024) return fail(h, ERTGPU_EINVAL, "chip length %d too large for the generic kernel", c.CL);
        CUDA_TRY(h, cudaFuncSetAttribute(demod_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((nblocks + nthr - 1) / nthr);
        demod_generic_kernel<<<grid, nthr, smem, st>>>(d_iq, hist, c.hist_samples, h->hist_valid, h->d_lut,
                                                       plane + c.hist_words, nblocks, c.BS, c.CL);
        CUDA_TRY(h, cudaGetLastError());
    }
    h->launches++;
    if (tm) CUDA_TRY(h, cudaEventRecord(h->ev_stage[1], st));

    // 2. preamble search over every start position of the call
    {
        SearchParams sp;
        SlideParams sl;
        if (!h->search_legacy && make_slide_params(c, p0, nwords, &sl)) {
            const int smem = 2 * sl.load_words * 4 + 16;
            const long long tiles = (nwords + sl.tile_words - 1) / sl.tile_words;
            int per_sm = (227 * 1024) / (smem + 1024);
            if (per_sm > 2048 / kSlideThreads) per_sm = 2048 / kSlideThreads;
            unsigned grid = (unsigned)std::min<long long>(tiles, (long long)h->sm_count * per_sm);
            if (grid < 1) grid = 1;
            const bool half = (c.SL % 32) == 16;
#define ERT_SLIDE_K(K)                                                                                           \
    do {                                                                                                         \
        static OncePerDevice once;                                                                               \
        if (once.first(h->device))                                                                               \
            CUDA_TRY(h, cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (2 * kSlideMaxLoad + 4))); \
        CUDA_TRY(h, launch_dep(pdl, K, dim3(grid), dim3(kSlideThreads), (size_t)smem, st, (const uint32_t*)plane, sl, h->d_hits, h->d_words, \
                               h->cand_cap, h->d_counters));                                                     \
    } while (0)
#define ERT_SLIDE(N, ...)                                                                    \
    do {                                                                                     \
        if (half) ERT_SLIDE_K((search_slide_kernel<N, true, ##__VA_ARGS__>));                \
        else ERT_SLIDE_K((search_slide_kernel<N, false, ##__VA_ARGS__>));                    \
    } while (0)
            if (c.npre == 1) {
                // single preamble: the stock protocols' first 16 bits are compile-time constants
                // (scm 0xF953, scm+ 0x16A3, idm/netidm 0x5555, r900 0x0000), anything else runs the generic probe
                switch (slide_pattern(c, 0)) {
                    case 0xF953u: ERT_SLIDE(1, 0xF953u); break;
                    case 0x16A3u: ERT_SLIDE(1, 0x16A3u); break;
                    case 0x5555u: ERT_SLIDE(1, 0x5555u); break;
                    case 0x0000u: ERT_SLIDE(1, 0x0000u); break;
                    default: ERT_SLIDE(1); break;
                }
            } else if (c.npre == 2) {
                ERT_SLIDE(2);
            } else if (c.npre == 3) {
                // scm, scm+, idm (or netidm) registered in that order: BASELINE.json configs[2]
                if (slide_pattern(c, 0) == 0xF953u && slide_pattern(c, 1) == 0x16A3u && slide_pattern(c, 2) == 0x5555u)
                    ERT_SLIDE(3, 0xF953u, 0x16A3u, 0x5555u);
                else
                    ERT_SLIDE(3);
            } else {
                // "-msgtype=all" = scm, scm+, idm, r900 (main.go:67-73)
                if (c.npre == 4 && slide_pattern(c, 0) == 0xF953u && slide_pattern(c, 1) == 0x16A3u &&
                    slide_pattern(c, 2) == 0x5555u && slide_pattern(c, 3) == 0x0000u)
                    ERT_SLIDE(4, 0xF953u, 0x16A3u, 0x5555u, 0x0000u);
                else
                    ERT_SLIDE(4);
            }
#undef ERT_SLIDE
#undef ERT_SLIDE_K
        } else if (make_search_params(c, p0, nwords, &sp)) {
            const size_t smem = 0;
            long long tiles = (nwords + kSearchTile - 1) / kSearchTile;
            unsigned grid = (unsigned)std::min<long long>(tiles, 148 * 4);
            if (grid < 1) grid = 1;
            const int mode = search_mode(c, p0);
#define ERT_SEARCH(N, M) CUDA_TRY(h, launch_dep(pdl, search_kernel<N, M>, dim3(grid), dim3(kSearchThreads), smem, st, (const uint32_t*)plane, sp, h->d_hits, h->d_words, h->cand_cap, h->d_counters))
#define ERT_SEARCH_N(N) do { if (mode == 1) ERT_SEARCH(N, 1); else if (mode == 2) ERT_SEARCH(N, 2); else ERT_SEARCH(N, 0); } while (0)
            switch (c.npre) {
                case 1: ERT_SEARCH_N(1); break;
                case 2: ERT_SEARCH_N(2); break;
                case 3: ERT_SEARCH_N(3); break;
                default: ERT_SEARCH_N(4); break;
            }
#undef ERT_SEARCH_N
#undef ERT_SEARCH
        } else {
            const int nthr = 256;
            long long blocks = std::min<long long>((nwords + nthr - 1) / nthr, 148 * 16);
            if (blocks < 1) blocks = 1;
            CUDA_TRY(h, launch_dep(pdl, search_generic_kernel, dim3((unsigned)blocks), dim3(nthr), 0, st, (const uint32_t*)plane, p0, nwords, c,
                                   h->d_hits, h->d_words, h->cand_cap, h->d_counters));
        }
        CUDA_TRY(h, cudaGetLastError());
        h->launches++;
    }

    if (tm) CUDA_TRY(h, cudaEventRecord(h->ev_stage[2], st));
    // 3. r900 payload digits for hits of an r900 preamble
    const uint8_t* digits = nullptr;
    if (h->has_r900) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_block_slot, 0xFF, (size_t)nblocks * sizeof(int), st));
        CUDA_TRY(h, cudaMemsetAsync(h->d_slot_count, 0, sizeof(unsigned int), st));
        r900_mark_kernel<<<148, 256, 0, st>>>(c, h->d_hits, h->cand_cap, h->d_counters, h->d_block_slot, h->d_slot_block,
                                              h->r900_slots, h->d_slot_count);
        CUDA_TRY(h, cudaGetLastError());
        // the two-warp pipeline loads 8 samples (16 bytes) at a time: chain starts and the call's bytes must be 16-byte aligned
        const bool vec_ok = (reinterpret_cast<uintptr_t>(d_iq) & 15u) == 0 && (reinterpret_cast<uintptr_t>(hist) & 15u) == 0 &&
                            c.BUF % 8 == 0 && c.BS % 8 == 0 && c.hist_samples % 8 == 0;
        if (h->r900_chain_mode == 2) {
            r900_chain_kernel<true><<<148 * 16, kR900ChainWarps * 32, 0, st>>>(d_iq, hist, c.hist_samples, h->hist_valid, h->d_lut, c,
                                                                       h->d_slot_block, h->r900_slots, h->d_slot_count,
                                                                       h->r900_span, h->d_r900_scratch);
        } else if (h->r900_chain_mode == 1 || !vec_ok) {
            r900_chain_kernel<false><<<148 * 16, kR900ChainWarps * 32, 0, st>>>(d_iq, hist, c.hist_samples, h->hist_valid, h->d_lut, c,
                                                                       h->d_slot_block, h->r900_slots, h->d_slot_count,
                                                                       h->r900_span, h->d_r900_scratch);
        } else if (h->r900_chain_mode == 3) {   // Tensor Memory between the stages (r900_tmem.cuh): exact, measured no faster (212 vs 203 us)
            static OncePerDevice t3_once;
            if (t3_once.first(current_device())) CUDA_TRY(h, cudaFuncSetAttribute(r900_chain3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3SmemBytes));
            r900_chain3_kernel<<<(unsigned)h->sm_count * 2, kT3Warps * 32, kT3SmemBytes, st>>>(d_iq, hist, c.hist_samples, h->hist_valid, c, h->d_slot_block,
                                                        h->r900_slots, h->d_slot_count, h->r900_span, h->d_r900_scratch);
        } else {   // default: producer warps + one consumer warp per 16 chains, hand-offs through shared memory
            r900_chain2_kernel<<<(unsigned)h->sm_count * 2, (kChainProducers + 1) * 32, 0, st>>>(d_iq, hist, c.hist_samples, h->hist_valid, c, h->d_slot_block,
                                                        h->r900_slots, h->d_slot_count, h->r900_span, h->d_r900_scratch);
        }
        CUDA_TRY(h, cudaGetLastError());
        r900_digits_kernel<<<148 * 16, 256, 0, st>>>(c, h->d_hits, h->cand_cap, h->d_counters, h->d_block_slot, h->r900_span,
                                                    h->d_r900_scratch, h->d_digits);
        CUDA_TRY(h, cudaGetLastError());
        // blocks that found no scratch slot: exact per-candidate replay
        r900_replay_kernel<<<148, 64, 0, st>>>(d_iq, hist, c.hist_samples, h->hist_valid, h->d_lut, c, h->d_hits,
                                               h->cand_cap, h->d_counters, h->d_block_slot, h->d_digits, h->d_slot_count, h->r900_slots);
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 4;
        digits = h->d_digits;
    }

    // 4. slice + integrity screens, one warp per word of starts with a hit; the same kernel carries the history to
    //    the other plane / history buffers (both only read what the earlier kernels wrote)
    {
        CarryArgs ca;
        ca.plane_src = plane; ca.plane_dst = plane_next; ca.src_offset_words = nwords; ca.hist_words = c.hist_words;
        ca.iq = d_iq; ca.hist_src = hist; ca.hist_dst = hist_next; ca.hist_samples = c.hist_samples;
        ca.nsamples = nblocks * c.BS;
        const unsigned grid = (unsigned)h->sm_count * 4;
        const bool dep = pdl && !h->has_r900;   // the r900 kernels between Search and Slice are ordinary launches
        if (c.PK <= 128)
            CUDA_TRY(h, launch_dep(dep, extract_words_kernel<4>, dim3(grid), dim3(kExtractWarps * 32), 0, st, (const uint32_t*)plane, p0, c,
                                   (const HitWord*)h->d_words, h->cand_cap, (const uint16_t*)h->d_crc, h->gf, digits, (long long)h->block_counter,
                                   flags, h->d_out, h->cand_cap, h->d_counters, ca));
        else
            CUDA_TRY(h, launch_dep(dep, extract_words_kernel<(ERTGPU_MAX_PACKET_BYTES * 8 + 31) / 32>, dim3(grid), dim3(kExtractWarps * 32), 0,
                                   st, (const uint32_t*)plane, p0, c, (const HitWord*)h->d_words, h->cand_cap, (const uint16_t*)h->d_crc, h->gf,
                                   digits, (long long)h->block_counter, flags, h->d_out, h->cand_cap, h->d_counters, ca));
        CUDA_TRY(h, cudaGetLastError());
        h->launches++;
    }

    if (tm) CUDA_TRY(h, cudaEventRecord(h->ev_stage[3], st));
    if (tm) CUDA_TRY(h, cudaEventRecord(h->ev_stage[4], st));
    h->stage_valid = tm;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_counters, h->d_counters, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));

    h->last_iq = d_iq;
    h->last_hist = hist;
    h->last_hist_valid = h->hist_valid;
    h->last_plane = plane;
    h->last_first_block = h->block_counter;
    h->last_nblocks = nblocks;
    h->pending = true;
    h->pending_stream = st;

    h->cur_plane ^= 1;
    h->cur_hist ^= 1;
    long long hv = (long long)h->hist_valid + nblocks * c.BS;
    h->hist_valid = (int)std::min<long long>(hv, c.hist_samples);
    h->block_counter += nblocks;
    return ERTGPU_OK;
}

bool cand_less(const ertgpu_candidate& a, const ertgpu_candidate& b) {
    if (a.block != b.block) return a.block < b.block;
    if (a.preamble_id != b.preamble_id) return a.preamble_id < b.preamble_id;
    return a.idx < b.idx;
}

// Wait for the pending pipeline and read its counters (candidates stay on the device).
int collect_sync(ertgpu_handle* h) {
    if (!h->pending) return ERTGPU_OK;
    CUDA_TRY(h, cudaStreamSynchronize(h->pending_stream));
    h->pending = false;
    const unsigned long long nh = h->h_counters[0], no = h->h_counters[1], nv = h->h_counters[2];
    h->total_hits += (int64_t)nh;
    h->total_valid += (int64_t)nv;
    if (nh > h->cand_cap) {
        h->overflow = true;
        h->need = std::max(h->need, nh);
    }
    h->uncopied_n = std::min(no, h->cand_cap);
    h->uncopied = h->uncopied_n > 0;
    return ERTGPU_OK;  // stage times are folded lazily (fold_stage_times)
}

// ... and append its candidates to h->results.
int collect(ertgpu_handle* h) {
    int rc = collect_sync(h);
    if (rc) return rc;
    if (h->uncopied) {
        const size_t old = h->results.size();
        h->results.resize(old + h->uncopied_n);
        CUDA_TRY(h, cudaMemcpy(h->results.data() + old, h->d_out, h->uncopied_n * sizeof(ertgpu_candidate), cudaMemcpyDeviceToHost));
        h->uncopied = false;
        // Slots are handed out by atomics, so a pipeline's records arrive unordered.  Pipelines of one
        // call cover ascending, disjoint block ranges: sorting each segment as it arrives (while the next
        // chunk's H2D copy is in flight) leaves the whole list in (block, preamble, idx) order.
        std::sort(h->results.begin() + (std::ptrdiff_t)old, h->results.end(), cand_less);
    }
    return ERTGPU_OK;
}

int deliver(ertgpu_handle* h, ertgpu_candidate* out, size_t cap, size_t* n_out) {
    if (h->overflow) {
        if (n_out) *n_out = (size_t)h->need;
        return fail(h, ERTGPU_ECAPACITY, "internal candidate capacity %llu exceeded (%llu hits): allocate with a larger max_candidates",
                    h->cand_cap, h->need);
    }
    if (n_out) *n_out = h->results.size();  // already sorted segment by segment in collect()
    if (h->results.size() > cap) return fail(h, ERTGPU_ECAPACITY, "output array holds %zu candidates, %zu needed", cap, h->results.size());
    if (!h->results.empty() && out) memcpy(out, h->results.data(), h->results.size() * sizeof(ertgpu_candidate));
    return ERTGPU_OK;
}

void begin_call(ertgpu_handle* h) {
    h->results.clear();
    h->uncopied = false;
    h->overflow = false;
    h->need = 0;
    h->total_hits = h->total_valid = 0;
}

}  // namespace

extern "C" {

int ertgpu_abi_version(void) { return ERTGPU_ABI_VERSION; }

const char* ertgpu_last_error(const ertgpu_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ertgpu_create(ertgpu_handle** out) {
    if (!out) return ERTGPU_EINVAL;
    ertgpu_handle* h = new (std::nothrow) ertgpu_handle();
    if (!h) return ERTGPU_ENOMEM;
    *out = h;
    return ERTGPU_OK;
}

void ertgpu_destroy(ertgpu_handle* h) {
    if (!h) return;
    free_device(h);
    delete h;
}

int ertgpu_stock_protocol(const char* msgtype, int32_t chip_length, ertgpu_protocol* p) {
    if (!msgtype || !p) return ERTGPU_EINVAL;
    memset(p, 0, sizeof(*p));
    p->data_rate = 32768;
    p->chip_length = chip_length;
    p->center_freq = 912600155u;
    snprintf(p->name, sizeof(p->name), "%s", msgtype);
    if (!strcmp(msgtype, "scm")) {  // scm/scm.go:40-50
        snprintf(p->preamble, sizeof(p->preamble), "%s", "111110010101001100000");
        p->preamble_symbols = 21;
        p->packet_symbols = 96;
        p->check_kind = ERTGPU_CHECK_CRC16;
        p->crc_init = 0; p->crc_poly = 0x6F63; p->crc_residue = 0;
        p->crc_from = 2; p->crc_to = 12;
    } else if (!strcmp(msgtype, "scm+")) {  // scmplus/scmplus.go:47-57
        snprintf(p->preamble, sizeof(p->preamble), "%s", "0001011010100011");
        p->preamble_symbols = 16;
        p->packet_symbols = 128;
        p->check_kind = ERTGPU_CHECK_CRC16;
        p->crc_init = 0xFFFF; p->crc_poly = 0x1021; p->crc_residue = 0x1D0F;
        p->crc_from = 2; p->crc_to = 16;
    } else if (!strcmp(msgtype, "idm") || !strcmp(msgtype, "netidm")) {  // idm/idm.go:46-56, netidm/netidm.go:58-68
        snprintf(p->preamble, sizeof(p->preamble), "%s", "01010101010101010001011010100011");
        p->preamble_symbols = 32;
        p->packet_symbols = 92 * 8;
        p->check_kind = ERTGPU_CHECK_IDM;
        p->crc_init = 0xFFFF; p->crc_poly = 0x1021; p->crc_residue = 0x1D0F;
        p->crc_from = 4; p->crc_to = 92;
    } else if (!strcmp(msgtype, "r900") || !strcmp(msgtype, "r900bcd")) {  // r900/r900.go:54-65
        snprintf(p->preamble, sizeof(p->preamble), "%s", "00000000000000001110010101100100");
        p->preamble_symbols = 32;
        p->packet_symbols = 116;
        p->center_freq = 912380000u;
        p->check_kind = ERTGPU_CHECK_R900;
    } else {
        return ERTGPU_EINVAL;  // parse.go:49 "invalid message type"
    }
    return ERTGPU_OK;
}

int ertgpu_register_protocol(ertgpu_handle* h, const ertgpu_protocol* p) {
    if (!h || !p) return ERTGPU_EINVAL;
    if (h->allocated) return fail(h, ERTGPU_EINVAL, "register_protocol after allocate");
    if (h->protos.size() >= ERTGPU_MAX_PROTOCOLS) return fail(h, ERTGPU_EINVAL, "too many protocols");
    const size_t n = strnlen(p->preamble, sizeof(p->preamble));
    if (n == 0 || n > ERTGPU_MAX_PREAMBLE) return fail(h, ERTGPU_EINVAL, "bad preamble length %zu", n);
    for (size_t i = 0; i < n; i++)
        if (p->preamble[i] != '0' && p->preamble[i] != '1') return fail(h, ERTGPU_EINVAL, "preamble must be '0'/'1'");
    if (p->chip_length <= 0 || p->preamble_symbols <= 0 || p->packet_symbols <= 0)
        return fail(h, ERTGPU_EINVAL, "non-positive geometry");
    if ((p->packet_symbols + 7) / 8 > ERTGPU_MAX_PACKET_BYTES) return fail(h, ERTGPU_EINVAL, "packet too long");
    if (p->check_kind == ERTGPU_CHECK_CRC16 &&
        (p->crc_from < 0 || p->crc_to > (p->packet_symbols + 7) / 8 || p->crc_from >= p->crc_to))
        return fail(h, ERTGPU_EINVAL, "bad crc range");
    if (p->check_kind == ERTGPU_CHECK_IDM && p->packet_symbols < 92 * 8) return fail(h, ERTGPU_EINVAL, "idm screen needs 92 bytes");
    h->protos.push_back(*p);
    // decode.go:105-109
    ertgpu_decoder_config& c = h->cfg;
    c.center_freq = p->center_freq;
    c.data_rate = std::max(c.data_rate, p->data_rate);
    c.chip_length = std::max(c.chip_length, p->chip_length);
    c.preamble_symbols = std::max(c.preamble_symbols, p->preamble_symbols);
    c.packet_symbols = std::max(c.packet_symbols, p->packet_symbols);
    c.n_protocols = (int32_t)h->protos.size();
    return ERTGPU_OK;
}

static int allocate_impl(ertgpu_handle* h, int32_t device, int64_t max_blocks_per_call, int64_t max_candidates);

int ertgpu_allocate(ertgpu_handle* h, int32_t device, int64_t max_blocks_per_call, int64_t max_candidates) {
    if (!h) return ERTGPU_EINVAL;
    if (h->allocated) return fail(h, ERTGPU_EINVAL, "already allocated");
    const int rc = allocate_impl(h, device, max_blocks_per_call, max_candidates);
    if (rc != ERTGPU_OK) {
        // a half-built handle must not pass the `allocated` checks of decode/tap/reset: release whatever was
        // created (the message of the failure is kept) and leave the handle ready for another ertgpu_allocate
        const std::string msg = h->err;
        free_device(h);
        h->cfg.n_preambles = 0;
        h->err = msg;
    }
    return rc;
}

static int allocate_impl(ertgpu_handle* h, int32_t device, int64_t max_blocks_per_call, int64_t max_candidates) {
    if (h->protos.empty()) return fail(h, ERTGPU_EINVAL, "no protocol registered");

    // decode.go:131-141
    ertgpu_decoder_config& c = h->cfg;
    c.symbol_length = c.chip_length << 1;
    c.sample_rate = c.data_rate * c.chip_length;
    c.preamble_length = c.preamble_symbols * c.symbol_length;
    c.packet_length = c.packet_symbols * c.symbol_length;
    c.block_size = next_pow2(c.preamble_length);
    c.block_size2 = c.block_size << 1;
    c.buffer_length = c.packet_length + c.block_size;
    c.packet_bytes = (c.packet_symbols + 7) >> 3;
    if (c.block_size < 32) return fail(h, ERTGPU_EINVAL, "block size %d below 32", c.block_size);

    DevCfg& d = h->dcfg;
    memset(&d, 0, sizeof(d));
    d.CL = c.chip_length; d.SL = c.symbol_length; d.BS = c.block_size;
    d.PS = c.preamble_symbols; d.PK = c.packet_symbols;
    d.PL = c.preamble_length; d.PKL = c.packet_length; d.BUF = c.buffer_length;
    d.words_per_block = d.BS / 32;
    d.hist_words = ((d.PKL + 31) / 32 + 3) & ~3;  // multiple of 4: the call's first word is 16-byte aligned
    d.hist_samples = d.PKL;
    d.packet_bytes = c.packet_bytes;
    d.nproto = (int32_t)h->protos.size();
    std::vector<uint16_t> tables, pos_tables;
    h->has_r900 = false;
    for (int i = 0; i < d.nproto; i++) {
        const ertgpu_protocol& p = h->protos[i];
        const int nb = (int)strnlen(p.preamble, sizeof(p.preamble));
        int found = -1;
        for (int j = 0; j < d.npre; j++) {
            if (d.pre_nbits[j] != nb) continue;
            bool same = true;
            for (int k = 0; k < nb; k++) same = same && d.pre_bits[j][k] == (uint8_t)(p.preamble[k] == '1');
            if (same) found = j;
        }
        if (found < 0) {  // decode.go:121-124
            found = d.npre++;
            d.pre_nbits[found] = nb;
            for (int k = 0; k < nb; k++) d.pre_bits[found][k] = (uint8_t)(p.preamble[k] == '1');
        }
        DevProto& dp = d.proto[i];
        dp.preamble_id = found;
        dp.check_kind = p.check_kind;
        dp.packet_bytes = (p.packet_symbols + 7) >> 3;
        dp.crc_from = p.crc_from; dp.crc_to = p.crc_to;
        dp.crc_init = p.crc_init; dp.crc_residue = p.crc_residue;
        dp.table = i;
        tables.resize((size_t)(i + 1) * 256);
        make_crc_table(p.crc_poly, tables.data() + (size_t)i * 256);
        {
            const uint16_t* tb = tables.data() + (size_t)i * 256;
            auto add_pos = [&](int n, int32_t* base, int32_t* cnt, uint16_t* k) {
                *base = (int32_t)(pos_tables.size() / 256);
                *cnt = n;
                for (int pp = 0; pp < n; pp++)
                    for (int v = 0; v < 256; v++) {
                        uint16_t crc = tb[v];                       // byte v processed with a zero register
                        for (int z = 0; z < n - 1 - pp; z++) crc = (uint16_t)((crc << 8) ^ tb[crc >> 8]);
                        pos_tables.push_back(crc);
                    }
                uint16_t crc = p.crc_init;                          // CRC(init, n zero bytes)
                for (int z = 0; z < n; z++) crc = (uint16_t)((crc << 8) ^ tb[crc >> 8]);
                *k = crc;
            };
            if (p.check_kind == ERTGPU_CHECK_CRC16) add_pos(p.crc_to - p.crc_from, &dp.pos_base, &dp.pos_n, &dp.pos_k);
            if (p.check_kind == ERTGPU_CHECK_IDM) {
                dp.crc_from = 4;
                add_pos(88, &dp.pos_base, &dp.pos_n, &dp.pos_k);
                add_pos(6, &dp.pos2_base, &dp.pos2_n, &dp.pos2_k);
            }
        }
        if (p.check_kind == ERTGPU_CHECK_R900) {
            d.pre_has_r900[found] = 1;
            h->has_r900 = true;
        }
    }
    c.n_preambles = d.npre;

    int ndev = 0;
    CUDA_TRY(h, cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(h, ERTGPU_ECUDA, "device %d not available (%d devices)", device, ndev);
    h->device = device;
    CUDA_TRY(h, cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(h, ERTGPU_ECUDA, "libertgpu is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);

    if (max_blocks_per_call <= 0) max_blocks_per_call = std::max<int64_t>(1, (64ll << 20) / c.block_size2);
    if (max_candidates <= 0) max_candidates = 1 << 18;
    h->max_blocks = max_blocks_per_call;
    h->cand_cap = (unsigned long long)max_candidates;
    h->allocated = true;  // from here free_device() cleans up on failure

    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_h2d[k], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
    }
    for (auto& set : h->ev_pool) for (int k = 0; k < 5; k++) CUDA_TRY(h, cudaEventCreate(&set[k]));
    h->ev_stage = h->ev_pool[0];
    (void)dynamic_smem_base();  // asked once, before the first launch needs it
    make_maglut(h->h_lut);
    make_gf32(&h->gf);
    CUDA_TRY(h, cudaMalloc(&h->d_lut, 256 * sizeof(float)));
    CUDA_TRY(h, cudaMemcpy(h->d_lut, h->h_lut, 256 * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMalloc(&h->d_crc, tables.size() * sizeof(uint16_t)));
    CUDA_TRY(h, cudaMemcpy(h->d_crc, tables.data(), tables.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMalloc(&h->d_crc_pos, std::max<size_t>(pos_tables.size(), 256) * sizeof(uint16_t)));
    if (!pos_tables.empty())
        CUDA_TRY(h, cudaMemcpy(h->d_crc_pos, pos_tables.data(), pos_tables.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));

    h->plane_words = (size_t)d.hist_words + (size_t)max_blocks_per_call * d.words_per_block + kSearchTile + kSearchMaxHalo + kSlideMaxLoad + 8;
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(h, cudaMalloc(&h->d_plane[k], h->plane_words * sizeof(uint32_t)));
        CUDA_TRY(h, cudaMemset(h->d_plane[k], 0, h->plane_words * sizeof(uint32_t)));
        CUDA_TRY(h, cudaMalloc(&h->d_hist[k], (size_t)d.hist_samples * 2 + 16));
        CUDA_TRY(h, cudaMemset(h->d_hist[k], 0, (size_t)d.hist_samples * 2 + 16));
    }
    CUDA_TRY(h, cudaMalloc(&h->d_hits, h->cand_cap * sizeof(RawHit)));
    CUDA_TRY(h, cudaMalloc(&h->d_words, h->cand_cap * sizeof(HitWord)));
    CUDA_TRY(h, cudaMalloc(&h->d_out, h->cand_cap * sizeof(ertgpu_candidate)));
    if (h->has_r900) {
        CUDA_TRY(h, cudaMalloc(&h->d_digits, h->cand_cap * ERTGPU_R900_DIGITS));
        h->r900_span = d.BS + d.PL - d.SL + 4 * ERTGPU_R900_DIGITS * d.CL + 1;
        if (h->r900_span > d.BUF + 1) h->r900_span = d.BUF + 1;
        // scratch slots: one per block that holds an r900 hit.  Sized for one packet per 16 blocks
        // (at least 64 slots), at most 1 GiB; blocks beyond that use the per-candidate replay.
        {
            const int64_t per_slot = (int64_t)h->r900_span * 4;
            int64_t want = std::max<int64_t>(64, max_blocks_per_call / 16);
            want = std::min<int64_t>(want, (1ll << 30) / per_slot);
            h->r900_slots = (int)std::max<int64_t>(1, std::min<int64_t>(want, max_blocks_per_call));
        }
        if (const char* e = getenv("ERTGPU_R900_SLOTS")) h->r900_slots = std::max(1, atoi(e));  // test hook: force the replay fallback
        CUDA_TRY(h, cudaMalloc(&h->d_block_slot, (size_t)max_blocks_per_call * sizeof(int)));
        CUDA_TRY(h, cudaMalloc(&h->d_slot_block, (size_t)h->r900_slots * sizeof(int)));
        CUDA_TRY(h, cudaMalloc(&h->d_slot_count, sizeof(unsigned int)));
        CUDA_TRY(h, cudaMalloc(&h->d_r900_scratch, (size_t)h->r900_slots * r900_pitch(h->r900_span) * sizeof(float) + 64));
    }
    CUDA_TRY(h, cudaMalloc(&h->d_counters, kCntN * sizeof(unsigned long long)));
    CUDA_TRY(h, cudaHostAlloc(&h->h_counters, kCntN * sizeof(unsigned long long), cudaHostAllocDefault));
    h->tap_floats = (size_t)std::max(d.BS + d.SL, d.BUF) * 2 + 16;
    CUDA_TRY(h, cudaMalloc(&h->d_tap, h->tap_floats * sizeof(float)));

    h->demod_variant = demod_fast_variant(d.CL, d.BS);
    if (const char* e = getenv("ERTGPU_FAST_WARPS")) h->demod_warps = atoi(e);   // 100 * VAR + W, see launch_demod_fast
    if (const char* e = getenv("ERTGPU_SEARCH_LEGACY")) h->search_legacy = atoi(e) != 0;
    if (const char* e = getenv("ERTGPU_PDL")) h->use_pdl = atoi(e) != 0;
    if (const char* e = getenv("ERTGPU_R900_CHAIN")) h->r900_chain_mode = strcmp(e, "shfl") == 0 ? 2 : (strcmp(e, "smem") == 0 ? 1 : (strcmp(e, "tmem") == 0 ? 3 : 0));
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (h->sm_count < 1) h->sm_count = 148;
    h->cur_plane = h->cur_hist = 0;
    h->hist_valid = 0;
    h->block_counter = 0;
    return ERTGPU_OK;
}

int ertgpu_get_config(const ertgpu_handle* h, ertgpu_decoder_config* cfg) {
    if (!h || !cfg) return ERTGPU_EINVAL;
    if (!h->allocated) return ERTGPU_EINVAL;
    *cfg = h->cfg;
    return ERTGPU_OK;
}

int ertgpu_reset(ertgpu_handle* h) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        CUDA_TRY(h, cudaStreamSynchronize(h->pending_stream));
        h->pending = false;
    }
    h->uncopied = false;
    // the history in front of the next call is zeroed on that call's stream (enqueue_pipeline)
    h->fresh = true;
    h->hist_valid = 0;
    h->block_counter = 0;
    begin_call(h);
    return ERTGPU_OK;
}

int ertgpu_set_demod_variant(ertgpu_handle* h, int32_t variant) {
    // test hook: 0 forces the generic kernel, -1 restores the automatic choice
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    if (variant < 0) h->demod_variant = demod_fast_variant(h->dcfg.CL, h->dcfg.BS);
    else if (variant == 0) h->demod_variant = 0;
    else return ERTGPU_EINVAL;
    return ERTGPU_OK;
}

int ertgpu_decode_device_async(ertgpu_handle* h, const void* d_iq, size_t nbytes, uint32_t flags, void* stream) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    if (!d_iq && nbytes) return fail(h, ERTGPU_EINVAL, "null input");
    if ((reinterpret_cast<uintptr_t>(d_iq) & 15) != 0) return fail(h, ERTGPU_EINVAL, "device input must be 16-byte aligned");
    const size_t bs2 = (size_t)h->cfg.block_size2;
    if (nbytes % bs2) return fail(h, ERTGPU_ESIZE, "nbytes %zu is not a multiple of BlockSize2 %zu", nbytes, bs2);
    const int64_t nblocks = (int64_t)(nbytes / bs2);
    if (nblocks > h->max_blocks) return fail(h, ERTGPU_ESIZE, "%lld blocks exceed max_blocks_per_call %lld", (long long)nblocks, (long long)h->max_blocks);
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        int rc = collect(h);
        if (rc) return rc;
    }
    begin_call(h);
    if (nblocks == 0) return ERTGPU_OK;
    return enqueue_pipeline(h, static_cast<const uint8_t*>(d_iq), nblocks, flags, stream ? (cudaStream_t)stream : h->stream);
}

int ertgpu_fetch(ertgpu_handle* h, ertgpu_candidate* out, size_t cap, size_t* n_out) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = collect(h);
    if (rc) return rc;
    return deliver(h, out, cap, n_out);
}

int ertgpu_decode(ertgpu_handle* h, const uint8_t* iq, size_t nbytes, uint32_t flags, ertgpu_candidate* out,
                  size_t cap, size_t* n_out) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    if (!iq && nbytes) return fail(h, ERTGPU_EINVAL, "null input");
    const size_t bs2 = (size_t)h->cfg.block_size2;
    if (nbytes % bs2) return fail(h, ERTGPU_ESIZE, "nbytes %zu is not a multiple of BlockSize2 %zu", nbytes, bs2);
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        int rc = collect(h);
        if (rc) return rc;
    }
    begin_call(h);
    const int64_t nblocks = (int64_t)(nbytes / bs2);
    if (nblocks == 0) return deliver(h, out, cap, n_out);

    // staging chunks (allocated on first use): H2D of chunk i+1 overlaps the kernels of chunk i
    int64_t chunk_mib = 32;
    if (const char* e = getenv("ERTGPU_CHUNK_MIB")) chunk_mib = std::max(1, atoi(e));  // tuning knob
    const int64_t chunk_blocks = std::min<int64_t>(h->max_blocks, std::max<int64_t>(1, (chunk_mib << 20) / (int64_t)bs2));
    if (h->stage_bytes < (size_t)chunk_blocks * bs2) {
        for (int k = 0; k < 2; k++) {
            cudaFree(h->d_stage[k]);
            h->d_stage[k] = nullptr;
        }
        h->stage_bytes = 0;
        for (int k = 0; k < 2; k++) CUDA_TRY(h, cudaMalloc(&h->d_stage[k], (size_t)chunk_blocks * bs2));
        h->stage_bytes = (size_t)chunk_blocks * bs2;
    }
    // Pinned (or registered / managed) input goes to the device directly.  Ordinary pageable memory -- a Go slice,
    // main.go:166's block buffer -- would make every cudaMemcpyAsync a synchronous driver-staged copy: it is staged
    // through two pinned buffers of the handle instead, filled by a few host threads while the previous chunk's
    // transfer is in flight.
    bool pageable = true;
    {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, iq) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (pageable && h->h_stage_bytes < h->stage_bytes) {
        for (int k = 0; k < 2; k++) {
            if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]);
            h->h_stage[k] = nullptr;
        }
        h->h_stage_bytes = 0;
        for (int k = 0; k < 2; k++) CUDA_TRY(h, cudaHostAlloc(&h->h_stage[k], h->stage_bytes, cudaHostAllocDefault));
        h->h_stage_bytes = h->stage_bytes;
    }
    if (pageable) {
        int copy_threads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 4));  // ~6 GB/s per memcpy stream; 8 fill a PCIe 5 x16 link
        if (const char* e = getenv("ERTGPU_STAGE_THREADS")) copy_threads = std::min(64, std::max(1, atoi(e)));
        if (!h->copy_pool || h->copy_pool->size() != copy_threads) h->copy_pool.reset(new CopyPool(copy_threads - 1));
    }
    int64_t done = 0, launches = 0;
    for (int64_t i = 0; done < nblocks; i++) {
        const int k = (int)(i & 1);
        const int64_t nb = std::min(chunk_blocks, nblocks - done);
        if (i >= 2) CUDA_TRY(h, cudaEventSynchronize(h->ev_done[k]));  // the kernels that read d_stage[k] (and its H2D) are done
        const uint8_t* src = iq + (size_t)done * bs2;
        const size_t nbytes_chunk = (size_t)nb * bs2;
        if (pageable) {
            h->copy_pool->copy(h->h_stage[k], src, nbytes_chunk);
            src = h->h_stage[k];
        }
        CUDA_TRY(h, cudaMemcpyAsync(h->d_stage[k], src, nbytes_chunk, cudaMemcpyHostToDevice, h->copy_stream));
        CUDA_TRY(h, cudaEventRecord(h->ev_h2d[k], h->copy_stream));
        if (i >= 1) {
            int rc = collect(h);
            if (rc) return rc;
        }
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_h2d[k], 0));
        int rc = enqueue_pipeline(h, h->d_stage[k], nb, flags, h->stream);
        if (rc) return rc;
        launches += h->launches;
        CUDA_TRY(h, cudaEventRecord(h->ev_done[k], h->stream));
        done += nb;
    }
    int rc = collect(h);
    if (rc) return rc;
    h->launches = launches;
    return deliver(h, out, cap, n_out);
}

int ertgpu_last_counts(ertgpu_handle* h, int64_t* n_candidates, int64_t* n_valid) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    if (h->pending) {
        CUDA_TRY(h, cudaSetDevice(h->device));
        int rc = collect_sync(h);  // counters only; ertgpu_fetch copies the candidates if wanted
        if (rc) return rc;
    }
    if (n_candidates) *n_candidates = h->total_hits;
    if (n_valid) *n_valid = h->total_valid;
    return ERTGPU_OK;
}

int64_t ertgpu_last_launches(const ertgpu_handle* h) { return h ? h->launches : 0; }

int ertgpu_tap(ertgpu_handle* h, int32_t which, int64_t block, void* dst, size_t cap, size_t* n_out) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        int rc = collect(h);
        if (rc) return rc;
    }
    if (!h->last_plane) return fail(h, ERTGPU_EINVAL, "no decode call yet");
    const DevCfg& c = h->dcfg;
    const int64_t b = block - h->last_first_block;
    if (b < 0 || b >= h->last_nblocks) return fail(h, ERTGPU_EINVAL, "block %lld outside the last (chunk of the) call [%lld,%lld)",
                                                   (long long)block, (long long)h->last_first_block,
                                                   (long long)(h->last_first_block + h->last_nblocks));
    std::vector<uint8_t> bytes;
    if (which == ERTGPU_TAP_SIGNAL || which == ERTGPU_TAP_CSUM) {
        float* sig = h->d_tap;
        float* cs = h->d_tap + (c.BS + c.SL);
        tap_signal_csum_kernel<<<1, 32, 0, h->stream>>>(h->last_iq, h->last_hist, c.hist_samples, h->last_hist_valid, h->d_lut,
                                                        b, c.BS, c.SL, sig, cs);
        CUDA_TRY(h, cudaGetLastError());
        const size_t n = which == ERTGPU_TAP_SIGNAL ? (size_t)(c.BS + c.SL) : (size_t)(c.BS + c.SL + 1);
        bytes.resize(n * sizeof(float));
        CUDA_TRY(h, cudaMemcpyAsync(bytes.data(), which == ERTGPU_TAP_SIGNAL ? sig : cs, bytes.size(), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else if (which == ERTGPU_TAP_QUANTIZED || which == ERTGPU_TAP_PACKED) {
        // Quantized after block b = stream bits [(b+1)*BS - BUF, (b+1)*BS)  (decode.go:166,172)
        const long long first_bit = (long long)c.hist_words * 32 + (b + 1) * c.BS - c.BUF;
        const long long w0 = first_bit >> 5, w1 = (first_bit + c.BUF + 31) >> 5;
        std::vector<uint32_t> words((size_t)(w1 - w0));
        CUDA_TRY(h, cudaMemcpy(words.data(), h->last_plane + w0, words.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        std::vector<uint8_t> q((size_t)c.BUF);
        for (long long i = 0; i < c.BUF; i++) {
            const long long pos = first_bit + i - (w0 << 5);
            q[(size_t)i] = (uint8_t)((words[(size_t)(pos >> 5)] >> (31 - (pos & 31))) & 1u);
        }
        if (which == ERTGPU_TAP_QUANTIZED) {
            bytes = q;
        } else {  // decode.go:259-265
            bytes.resize((size_t)((c.BS + c.PL + 7) >> 3));
            for (size_t B = 0; B < bytes.size(); B++) {
                uint8_t v = 0;
                for (int k = 0; k < 8; k++) v = (uint8_t)((v << 1) | q[B * 8 + k]);
                bytes[B] = v;
            }
        }
    } else if (which == ERTGPU_TAP_R900_QUANTIZED) {
        float* cs = h->d_tap;
        uint8_t* qd = reinterpret_cast<uint8_t*>(h->d_tap + c.BUF + 4);
        tap_r900_csum_kernel<<<1, 32, 0, h->stream>>>(h->last_iq, h->last_hist, c.hist_samples, h->last_hist_valid, h->d_lut, b,
                                                      c.BS, c.BUF, cs);
        CUDA_TRY(h, cudaGetLastError());
        tap_r900_digits_kernel<<<(c.BUF + 255) / 256, 256, 0, h->stream>>>(cs, c.BUF, c.CL, qd);
        CUDA_TRY(h, cudaGetLastError());
        bytes.resize((size_t)c.BUF);
        CUDA_TRY(h, cudaMemcpyAsync(bytes.data(), qd, bytes.size(), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else {
        return fail(h, ERTGPU_EINVAL, "unknown tap %d", which);
    }
    if (n_out) *n_out = bytes.size();
    if (dst) memcpy(dst, bytes.data(), std::min(cap, bytes.size()));
    return ERTGPU_OK;
}

int ertgpu_set_stage_timing(ertgpu_handle* h, int32_t enable) {
    if (!h || !h->allocated) return ERTGPU_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        int rc = collect_sync(h);
        if (rc) return rc;
    }
    h->stage_timing = enable != 0;
    if (enable) {
        for (int k = 0; k < 4; k++) h->stage_sum[k] = 0;
        h->stage_n = 0;
        h->stage_unread = 0;
    }
    return ERTGPU_OK;
}

int ertgpu_last_stage_ms(ertgpu_handle* h, float* ms4) {
    if (!h || !h->allocated || !ms4) return ERTGPU_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        int rc = collect_sync(h);
        if (rc) return rc;
    }
    if (!h->stage_valid) return fail(h, ERTGPU_EINVAL, "stage timing was not enabled for the last decode");
    for (int k = 0; k < 4; k++) CUDA_TRY(h, cudaEventElapsedTime(&ms4[k], h->ev_stage[k], h->ev_stage[k + 1]));
    return ERTGPU_OK;
}

int ertgpu_stage_ms_mean(ertgpu_handle* h, float* ms4, int64_t* n_pipelines) {
    if (!h || !h->allocated || !ms4) return ERTGPU_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->pending) {
        int rc = collect_sync(h);
        if (rc) return rc;
    }
    fold_stage_times(h);
    for (int k = 0; k < 4; k++) ms4[k] = h->stage_n ? (float)(h->stage_sum[k] / (double)h->stage_n) : 0.0f;
    if (n_pipelines) *n_pipelines = h->stage_n;
    return ERTGPU_OK;
}

// "0-31,64-95" -> cpu_set_t; returns the number of CPUs
static int parse_cpulist(const char* str, cpu_set_t* set) {
    CPU_ZERO(set);
    int n = 0;
    const char* p = str;
    while (*p) {
        char* end = nullptr;
        long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        p = end;
        if (*p == '-') {
            b = strtol(p + 1, &end, 10);
            if (end == p + 1) break;
            p = end;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++)
            if (c >= 0 && !CPU_ISSET((int)c, set)) { CPU_SET((int)c, set); n++; }
        while (*p == ',' || *p == ' ' || *p == '\n') p++;
    }
    return n;
}

int ertgpu_bind_host_thread(int32_t device, int32_t* ncpus, int32_t* numa_node) {
    if (ncpus) *ncpus = 0;
    if (numa_node) *numa_node = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return ERTGPU_ECUDA;
    }
    for (char* c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
    char path[128], buf[4096] = {0};
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    if (FILE* f = fopen(path, "r")) {
        int node = -1;
        if (fscanf(f, "%d", &node) == 1 && numa_node) *numa_node = node;
        fclose(f);
    }
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return ERTGPU_EINVAL;
    const size_t got = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    buf[got] = 0;
    cpu_set_t want, allowed, both;
    if (parse_cpulist(buf, &want) <= 0) return ERTGPU_EINVAL;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return ERTGPU_EINVAL;
    CPU_AND(&both, &want, &allowed);
    const int n = CPU_COUNT(&both);
    if (n <= 0) return ERTGPU_EINVAL;          // the GPU's CPUs are outside this process's cpuset: leave it alone
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return ERTGPU_EINVAL;
    if (ncpus) *ncpus = n;
    return ERTGPU_OK;
}

const char* ertgpu_last_kernels(const ertgpu_handle* h) { return h ? h->kernels.c_str() : ""; }

int ertgpu_host_alloc(void** out, size_t nbytes) {
    if (!out) return ERTGPU_EINVAL;
    cudaError_t e = cudaHostAlloc(out, nbytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? ERTGPU_ENOMEM : ERTGPU_ECUDA;
    }
    return ERTGPU_OK;
}

int ertgpu_host_free(void* p) {
    if (!p) return ERTGPU_OK;
    return cudaFreeHost(p) == cudaSuccess ? ERTGPU_OK : ERTGPU_ECUDA;
}

int ertgpu_synth_fill(int32_t device, void* d_out, int64_t first_sample, int64_t nsamples, uint64_t seed,
                      const ertgpu_synth_packet* packets, int64_t npackets, void* stream) {
    if (!d_out || nsamples < 0 || npackets < 0 || (npackets && !packets)) return ERTGPU_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return ERTGPU_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    ertgpu_synth_packet* d_pk = nullptr;
    // only the packets that can touch [first_sample, first_sample + nsamples) are uploaded
    int64_t lo = 0, hi = npackets;
    while (lo < npackets && packets[lo].start_sample + (int64_t)packets[lo].n_chips * packets[lo].chip_length <= first_sample) lo++;
    while (hi > lo && packets[hi - 1].start_sample >= first_sample + nsamples) hi--;
    const int64_t n = hi - lo;
    if (cudaMalloc(&d_pk, std::max<int64_t>(n, 1) * sizeof(ertgpu_synth_packet)) != cudaSuccess) return ERTGPU_ENOMEM;
    if (n && cudaMemcpyAsync(d_pk, packets + lo, (size_t)n * sizeof(ertgpu_synth_packet), cudaMemcpyHostToDevice, st) != cudaSuccess) {
        cudaFree(d_pk);
        return ERTGPU_ECUDA;
    }
    long long groups = (nsamples + 7) / 8;
    long long blocks = std::min<long long>((groups + 255) / 256, 148 * 32);
    if (blocks < 1) blocks = 1;
    synth_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<uint8_t*>(d_out), first_sample, nsamples, seed, d_pk, n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_pk);
    return e == cudaSuccess ? ERTGPU_OK : ERTGPU_ECUDA;
}

}  // extern "C"
