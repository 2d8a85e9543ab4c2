// demod_fast.cuh -- the hot kernel: magnitude + Manchester matched filter + quantize + pack with
// the chip length as a compile-time constant.
//
// Replaces, fused into one pass over the IQ bytes: MagLUT.Execute (protocol/decode.go:219-225),
// Filter (decode.go:229-245) and the bit packing of Search (decode.go:259-265).
//
// Design (see DESIGN.md "demod_fast"):
//  * One thread per reference block ("chain"); a warp owns 32 consecutive blocks.  The float32
//    running sum of Filter is sequential within a block and restarts at every block
//    (decode.go:232-236), so this is the only parallelism that keeps the sign bits exact.
//  * Per sample: csum c += m; chip sum A = c - c[CL ago]; filter f = A[CL ago] - A; the two
//    histories are rings of CL values each held in REGISTERS: the step loop is unrolled CL times
//    so every ring slot is a fixed register (no indexing, no shared-memory traffic).
//  * Magnitude lut[I] + lut[Q]: the 256-entry table is replicated per lane in shared memory as
//    [value][lane] with a 256-byte row pitch, so a lookup address is ONE byte-permute
//    (PRMT: {lane*4, value, base>>16}) and the 32 lanes always hit 32 different banks.  Column 32
//    of every row holds 0.0f: the first block of a stream reads its SL lead-in magnitudes there
//    (the reference's Signal buffer starts as zeros, decode.go:144).
//  * IQ bytes reach the thread that needs them through shared memory: each lane issues one
//    cp.async.bulk (TMA bulk copy, 2*CL bytes = one unrolled loop body's worth of its own chain)
//    per body into a 2-stage ring, completion on an mbarrier; rows are padded to an odd number
//    of 16-byte units so the per-lane LDS.128 reads are bank-conflict free.  HBM is read in
//    whole 32-byte sectors, each byte once.
//  * Sign bits are shifted into words with one funnel shift per sample and appended to the
//    per-block bit-plane row.
#pragma once

#include <cuda.h>

#include "ert_common.cuh"

namespace ert {

constexpr int kFastWarpChains = 32;
constexpr int kLutBytes = 65536;  // 256 rows x 256 B (lanes 0..31 + zero column at byte 128)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// 2D tiled TMA load: box (x = byte column, y = block row) -> dense smem tile
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// LUT read: not volatile, the table is constant after the prologue barrier
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <int CL>
struct FastGeom {
    static constexpr int kRowBytes = 2 * CL;                       // one body's IQ bytes per chain
    static constexpr int kRowUnits = kRowBytes / 16;
    static constexpr int kStrideUnits = kRowUnits | 1;             // odd pitch in 16-byte units: conflict-free LDS.128
    static constexpr int kStrideBytes = kStrideUnits * 16;
    static constexpr int kStageBytes = 32 * kStrideBytes;
    static constexpr int kWarpBytes = 2 * kStageBytes;             // 2-stage ring
    static constexpr int kFullWords = CL / 32;                     // whole output words per body
    static constexpr int kTailBits = CL % 32;
    static_assert(CL % 8 == 0, "fast kernel needs CL % 8 == 0 (16-byte rows, 8 samples per LDS.128)");
};

// smem map (byte offsets inside the dynamic segment, computed at run time):
//   [bars: 2 x 8 B per warp][pad][staging of the first warps ...][LUT at the next 64 KiB boundary]
//   [staging of the remaining warps ...]
template <int CL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
demod_fast_kernel(const __grid_constant__ CUtensorMap iq_map, const uint8_t* __restrict__ iq,
                  const uint8_t* __restrict__ hist, int hist_samples, int hist_valid,
                  const float* __restrict__ lut_g, uint32_t* __restrict__ plane_out, long long nblocks, int BS,
                  unsigned long long* __restrict__ tile_counter) {
    using G = FastGeom<CL>;
    extern __shared__ __align__(128) uint8_t fast_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sbase = smem_u32(fast_smem);
    const uint32_t lut_base = (sbase + 1024 + 0xFFFFu) & ~0xFFFFu;  // 64 KiB aligned, after the barriers
    // staging areas: as many warps as fit below the LUT, the rest above it
    const uint32_t below0 = sbase + 1024;
    const int nbelow = (int)((lut_base - below0) / G::kWarpBytes);
    const uint32_t stage0 = (warp < nbelow) ? below0 + warp * G::kWarpBytes
                                            : lut_base + kLutBytes + (warp - nbelow) * G::kWarpBytes;
    const uint32_t bar0 = sbase + warp * 16;  // two mbarriers per warp

    // ---- prologue: LUT [v][lane] + zero column, barriers ----
    for (int i = threadIdx.x; i < 256 * 33; i += WARPS * 32) {
        const int v = i / 33, l = i % 33;
        const float x = (l < 32) ? lut_g[v] : 0.0f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(lut_base + v * 256 + l * 4), "f"(x) : "memory");
    }
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int SL = 2 * CL;
    const int wpb = BS >> 5;                              // words per block
    const int nbody = (SL + BS - 1 + CL - 1) / CL;        // bodies of CL steps covering steps 0..SL+BS-2
    const int chain_units = (2 * (SL + BS)) >> 4;         // lead-in + block, in 16-byte units
    const long long ntiles = (nblocks + 31) / 32;
    const uint32_t lo_main = lut_base | (uint32_t)(lane * 4);
    const uint32_t row = stage0 + lane * G::kStrideBytes;
    const bool have_hist = hist_valid >= SL;
    const uint8_t* hist_lead = hist + 2ll * (hist_samples - SL);
    uint32_t phases = 0;  // bit s = parity to wait for on stage s

    for (;;) {
        unsigned long long tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1ull);
        tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
        if ((long long)tile >= ntiles) break;

        long long b = (long long)tile * 32 + lane;
        const bool live = b < nblocks;
        if (!live) b = nblocks - 1;  // duplicate a valid chain, results discarded
        // The chain's bytes: bodies 0,1 are the SL lead-in samples (the tail of the previous
        // block), bodies >= 2 the block itself -- one contiguous range except for block 0 of the
        // call, whose lead-in lives in the history buffer (or is all-zero magnitude at stream start).
        const uint32_t blk16 = (uint32_t)((b * 2ll * BS) >> 4);  // block start in 16-byte units
        const bool first = (b == 0);
        const uint32_t lo_lead = (first && !have_hist) ? (lut_base | 128u) : lo_main;
        // Body t of the 32 chains of this work tile = one [32 rows][2*CL bytes] box of the IQ
        // matrix (row = reference block, BlockSize2 bytes per row): a single TMA tile load issued
        // by one lane.  Rows past the end of the call and columns past the end of a row are
        // zero-filled by the TMA unit (the steps that read them produce no output).
        // The first work tile of a call is the exception: block 0 takes its lead-in from the
        // history buffer, so that tile uses one 1D bulk copy per lane instead.
        const bool lane_copies = (tile == 0);
        auto issue = [&](int t) {
            if (t >= nbody) return;
            const uint32_t bar = bar0 + (t & 1) * 8;
            if (!lane_copies) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(bar, 32u * G::kStrideBytes);
                    const int x = (t < 2) ? 2 * BS - 2 * G::kRowBytes + t * G::kRowBytes : (t - 2) * G::kRowBytes;
                    const int y = (int)(tile * 32) - (t < 2 ? 1 : 0);
                    tma_load_2d(stage0 + (t & 1) * G::kStageBytes, &iq_map, x, y, bar);
                }
                return;
            }
            const int off16 = t * G::kRowUnits;
            int n16 = chain_units - off16;
            if (n16 > G::kRowUnits) n16 = G::kRowUnits;
            const uint8_t* src;
            if (t < 2 && first) src = have_hist ? hist_lead + 16ll * off16 : iq;
            else src = iq + 16ll * ((long long)blk16 + off16 - 2 * G::kRowUnits);
            if (lane == 0) mbar_arrive_expect_tx(bar, 32u * (uint32_t)n16 * 16u);
            __syncwarp();
            bulk_g2s(row + (t & 1) * G::kStageBytes, src, (uint32_t)n16 * 16u, bar);
        };
        issue(0);
        issue(1);

        float cr[CL], ar[CL];
#pragma unroll
        for (int j = 0; j < CL; j++) { cr[j] = 0.0f; ar[j] = 0.0f; }
        float c = 0.0f;
        uint32_t acc = 0;   // pending output bits (low nacc bits), still as SIGN bits (inverted at store)
        int nacc = 0;       // warp-uniform
        int wi = 0;

        for (int t = 0; t < nbody; t++) {
            const int st = t & 1;
            mbar_wait(bar0 + st * 8, (phases >> st) & 1u);
            phases ^= 1u << st;
            const uint32_t lo = (t < 2) ? lo_lead : lo_main;
            const uint32_t src = row + st * G::kStageBytes;
            const bool emit = t >= 2;
            uint32_t w = 0;

#pragma unroll
            for (int g = 0; g < CL / 8; g++) {
                const uint4 v = lds128(src + g * 16);
                const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        const int j = g * 8 + q * 2 + s;
                        // {byte0: lane*4 (or the zero column), byte1: I or Q, bytes 2-3: LUT base >> 16}
                        const uint32_t ai = __byte_perm(xs[q], lo, s ? 0x7624 : 0x7604);
                        const uint32_t aq = __byte_perm(xs[q], lo, s ? 0x7634 : 0x7614);
                        const float m = __fadd_rn(lds_f32(ai), lds_f32(aq));   // decode.go:222
                        c = __fadd_rn(c, m);                                   // csum[k+1], decode.go:234
                        const float a = __fsub_rn(c, cr[j]);                   // csum[k+1] - csum[k+1-CL]
                        cr[j] = c;
                        const float f = __fsub_rn(ar[j], a);                   // decode.go:242
                        ar[j] = a;
                        w = __funnelshift_l(__float_as_uint(f), w, 1);         // sign bit in
                        if ((j & 31) == 31) {
                            // 32 more bits complete: emit one word (acc keeps the nacc pending bits)
                            if (emit) {
                                const uint32_t o = __funnelshift_r(w, acc, nacc);
                                if (live && wi < wpb) plane_out[b * wpb + wi] = ~o;
                                wi++;
                                acc = w;
                            }
                        }
                    }
                }
            }
            // every lane has read its row of this stage into registers: refill it with body t+2
            __syncwarp();
            issue(t + 2);

            if constexpr (G::kTailBits != 0) {
                // branch-free append of the body's last CL%32 bits (state only advances when emit)
                const uint32_t y = w & ((1u << G::kTailBits) - 1u);
                const uint32_t hi = acc >> (32 - G::kTailBits);
                const uint32_t lw = (acc << G::kTailBits) | y;
                const int n2 = nacc + G::kTailBits;
                const bool full = emit && n2 >= 32;
                const uint32_t o = __funnelshift_r(lw, hi, n2 & 31);
                if (full && live && wi < wpb) plane_out[b * wpb + wi] = ~o;
                wi += full ? 1 : 0;
                nacc = emit ? (n2 & 31) : nacc;
                acc = emit ? lw : acc;
            }
            if (t == 1) {  // body 1 contributes only its last step = f[0]
                acc = w & 1u;
                nacc = 1;
            }
        }
        __syncwarp();
    }
}

// ---- host side -------------------------------------------------------------------------

template <int CL>
constexpr int fast_warps() {
    // registers: 2*CL ring values + ~30 working registers; the register file is handed out in
    // units that make 8/12/16 warps per SM the useful steps (255/168/128 registers per thread)
    return (2 * CL + 40 <= 128) ? 16 : ((2 * CL + 40 <= 168) ? 12 : 8);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int CL, int W>
int launch_demod_fast_cw(const uint8_t* iq, const uint8_t* hist, int hist_samples, int hist_valid,
                         const float* lut, uint32_t* plane_out, long long nblocks, int BS,
                         unsigned long long* tile_counter, cudaStream_t st) {
    using G = FastGeom<CL>;
    auto kern = demod_fast_kernel<CL, W>;
    const int smem = 227 * 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    // the IQ bytes of the call as a 2D uint8 tensor: [nblocks rows][BlockSize2 bytes]
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)(2 * BS), (cuuint64_t)nblocks};
    const cuuint64_t gstride[1] = {(cuuint64_t)(2 * BS)};
    const cuuint32_t box[2] = {(cuuint32_t)G::kStrideBytes, 32u};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(iq), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    const long long ntiles = (nblocks + 31) / 32;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (ntiles + W - 1) / W;
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, W * 32, smem, st>>>(map, iq, hist, hist_samples, hist_valid, lut, plane_out, nblocks, BS, tile_counter);
    return (int)cudaGetLastError();
}

// chip lengths with a specialised kernel: every value the reference CLI accepts (flags.go:127-132
// allows 8,32,40,...,96) except 8 (BlockSize 512: rows of 16 bytes are too small to stage);
// anything else (e.g. 78, the fixture's capture rate) uses the generic kernel.
inline int demod_fast_variant(int CL, int BS) {
    (void)BS;
    switch (CL) {
        case 32: case 40: case 48: case 56: case 64: case 72: case 80: case 88: case 96: return CL;
        default: return 0;
    }
}

// warps == 0 picks the default resident-warp count for the chip length
inline int launch_demod_fast(int variant, int warps, const uint8_t* iq, const uint8_t* hist, int hist_samples,
                             int hist_valid, const float* lut, uint32_t* plane_out, long long nblocks,
                             int BS, unsigned long long* tile_counter, cudaStream_t st) {
#define ERT_FAST_ARGS iq, hist, hist_samples, hist_valid, lut, plane_out, nblocks, BS, tile_counter, st
#define ERT_FAST_CASE(N) \
    case N: return launch_demod_fast_cw<N, fast_warps<N>()>(ERT_FAST_ARGS);
    if (variant == 72 && warps == 8) return launch_demod_fast_cw<72, 8>(ERT_FAST_ARGS);
    if (variant == 72 && warps == 12) return launch_demod_fast_cw<72, 12>(ERT_FAST_ARGS);
    switch (variant) {
        ERT_FAST_CASE(32) ERT_FAST_CASE(40) ERT_FAST_CASE(48) ERT_FAST_CASE(56) ERT_FAST_CASE(64)
        ERT_FAST_CASE(72) ERT_FAST_CASE(80) ERT_FAST_CASE(88) ERT_FAST_CASE(96)
        default: return (int)cudaErrorInvalidValue;
    }
#undef ERT_FAST_CASE
#undef ERT_FAST_ARGS
}

}  // namespace ert
