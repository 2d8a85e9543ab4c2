// r900_tmem.cuh -- the r900 parser's running sum (r900/r900.go:96-100) with Tensor Memory as the mailbox between warps.
//
// One chain = ~24 700 DEPENDENT float32 adds (4.2 cycles each on B200), and the kernel's time is one chain's time.  What
// the warp that runs the adds pays on top is every OTHER instruction in its in-order stream: measured, a shared- or
// global-memory instruction between two dependent adds costs the chain 5-12 cycles (r900_chain2_kernel: one LDS.128 and
// one STG.128 per four adds = 15 cycles per add instead of 4.2).  Tensor Memory moves 32 values per lane with ONE
// instruction (tcgen05.ld/st.32x32b.x32), and a warp's lane i reads exactly what lane i of another warp of the same lane
// quarter wrote.  So the warp that adds touches memory twice per 32 adds:
//
//   producers (8 warps)   asynchronous 16-byte copies of the chains' IQ bytes, magnitudes computed (no table), STS -> ring
//   stager    (1 warp)   lane = chain: 8 LDS.128 -> tcgen05.st.x32 into the IN ring of Tensor Memory
//   adder     (1 warp)   tcgen05.ld.x32 -> 32 dependent adds (strictly left to right, r900.go:97-99) -> tcgen05.st.x32
//   writer    (1 warp)   tcgen05.ld.x32 from the OUT ring -> 8 STG.128 per lane to the chain's scratch row
//
// stager, adder and writer are warps 4, 8 and 0 of the CTA: all of lane quarter 0 (warp % 4), which is what lets them see
// each other's Tensor Memory lanes.  32 chains per CTA (lane = chain), one CTA per SM (it owns the SM's 512 columns); mbarrier rings between the stages, tcgen05 fences around every hand-off.
// Same additions in the same order as the reference: identical sums (tests/test_gpu_parity.py, r900 digits + quantized tap).
#pragma once

#include "search.cuh"

namespace ert {

#define ERT_R32(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) \
                   X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31)

__device__ __forceinline__ void tm_ld32(uint32_t taddr, float (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]),
          "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]),
          "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tm_st32(uint32_t taddr, const float (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]), "f"(r[9]), "f"(r[10]),
          "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]), "f"(r[16]), "f"(r[17]), "f"(r[18]), "f"(r[19]), "f"(r[20]), "f"(r[21]),
          "f"(r[22]), "f"(r[23]), "f"(r[24]), "f"(r[25]), "f"(r[26]), "f"(r[27]), "f"(r[28]), "f"(r[29]), "f"(r[30]), "f"(r[31])
        : "memory");
}
// the 32 loaded registers may be read after this
__device__ __forceinline__ void tm_wait_ld32(float (&r)[32]) {
#define ERT_X(i) "+f"(r[i]),
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]), "+f"(r[9]),
                   "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]),
                   "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]),
                   "+f"(r[30]), "+f"(r[31])
                 :
                 : "memory");
#undef ERT_X
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

constexpr int kT3Chains = 16;                 // chains per CTA: every other lane of the quarter-0 warps (two CTAs per SM share its Tensor Memory:
                                              // 2400 chains of a 4 GiB call = 150 CTAs on 148 SMs, and the conversion work spreads over all of them)
constexpr int kT3LaneShift = 1;               // chain = lane >> shift
constexpr int kT3Producers = 4;                // 4 chains each
constexpr int kT3Warps = 12;                  // warps 8 adder, 4 stager, 0 writer (lane quarter 0); 8 of the others produce (384 threads: no register squeeze)
constexpr int kT3Unit = 64;                   // samples per hand-off
constexpr int kT3Ring = 3;                    // shared-memory ring (producers -> stager)
constexpr int kT3TRing = 2;                   // Tensor Memory rings (stager -> adder, adder -> writer): 2 x 64 columns each, 256 per CTA
constexpr int kT3Cols = 2 * kT3TRing * kT3Unit;
constexpr int kT3Ahead = 4;                   // asynchronous copies in flight per producer lane
constexpr int kT3Pitch = kT3Unit + 4;
constexpr int kT3ChainsPerWarp = kT3Chains / kT3Producers;
constexpr int kT3Lanes = 32 / kT3ChainsPerWarp;
constexpr int kT3Loads = kT3Unit / (8 * kT3Lanes);
static_assert(kT3Loads >= 1 && (kT3Unit / 32) % 2 == 0 && (kT3Cols == 256 || kT3Cols == 512), "geometry");
constexpr int kT3MsBytes = kT3Ring * kT3Chains * kT3Pitch * 4;
constexpr int kT3RawBytes = kT3Producers * kT3Ahead * kT3Loads * 32 * 16;
constexpr int kT3SmemBytes = kT3MsBytes + kT3RawBytes + (2 * kT3Ring + 4 * kT3TRing) * 8;

__global__ void __launch_bounds__(kT3Warps * 32)
r900_chain3_kernel(const uint8_t* __restrict__ iq, const uint8_t* __restrict__ hist, int hist_samples, int hist_valid, DevCfg cfg,
                   const int* __restrict__ slot_block, int slot_cap, const unsigned int* __restrict__ slot_count, int span,
                   float* __restrict__ scratch) {
    extern __shared__ __align__(16) uint8_t t3_smem[];
    float (*m_s)[kT3Chains][kT3Pitch] = reinterpret_cast<float (*)[kT3Chains][kT3Pitch]>(t3_smem);
    uint4 (*raw_s)[kT3Ahead][kT3Loads][32] = reinterpret_cast<uint4 (*)[kT3Ahead][kT3Loads][32]>(t3_smem + kT3MsBytes);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(t3_smem + kT3MsBytes + kT3RawBytes);
    __shared__ uint32_t tbase_s;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = full0 + 8 * kT3Ring;                       // producers <-> stager
    const uint32_t tfull0 = empty0 + 8 * kT3Ring, tempty0 = tfull0 + 8 * kT3TRing;                // stager <-> adder
    const uint32_t ofull0 = tempty0 + 8 * kT3TRing, oempty0 = ofull0 + 8 * kT3TRing;              // adder <-> writer
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int r = 0; r < kT3Ring; r++) { mbar_init(full0 + 8 * r, kT3Producers); mbar_init(empty0 + 8 * r, 1); }
        for (int r = 0; r < kT3TRing; r++) {
            mbar_init(tfull0 + 8 * r, 1); mbar_init(tempty0 + 8 * r, 1);
            mbar_init(ofull0 + 8 * r, 1); mbar_init(oempty0 + 8 * r, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(kT3Cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tin0 = tbase_s, tout0 = tbase_s + kT3TRing * kT3Unit;   // lane quarter 0: lane field 0
    unsigned int nslots = *slot_count;
    if (nslots > (unsigned)slot_cap) nslots = (unsigned)slot_cap;
    const int nunits = (span - 1 + kT3Unit - 1) / kT3Unit;   // magnitudes 0 .. span-2 give csum[1 .. span-1]
    uint32_t ph = 0, ph2 = 0;   // parities of the barriers this role waits on
    const bool is_adder = warp == 8, is_stager = warp == 4, is_writer = warp == 0;   // the arbiter favours the highest warp id of a scheduler
    const int prod = warp - 1 - (warp > 4) - (warp > 8);   // rank among the other warps: 0..8, the first 8 produce

    for (unsigned int slot0 = blockIdx.x * kT3Chains; slot0 < nslots; slot0 += gridDim.x * kT3Chains) {
        if (is_adder) {
            // ---- adder: nothing in its instruction stream but the adds, one TMEM load and one TMEM store per 32 of them
            float acc = 0.0f;
            float va[32], vb[32];
            for (int u = 0; u < nunits; u++) {
                const int ts = u % kT3TRing;
                mbar_spin(tfull0 + 8 * ts, (ph >> ts) & 1u);
                ph ^= 1u << ts;
                if (u >= kT3TRing) {   // the writer has drained this OUT slot's previous unit
                    mbar_spin(oempty0 + 8 * ts, (ph2 >> ts) & 1u);
                    ph2 ^= 1u << ts;
                }
                tm_fence_after();
                const uint32_t tin = tin0 + ts * kT3Unit, tout = tout0 + ts * kT3Unit;
                tm_ld32(tin, va);
                auto group = [&](int g, float (&cur)[32], float (&nxt)[32]) {
                    tm_wait_ld32(cur);
                    if (g + 1 < kT3Unit / 32) tm_ld32(tin + (g + 1) * 32, nxt);
#pragma unroll
                    for (int k = 0; k < 32; k++) {          // strictly left to right, r900.go:97-99
                        acc = __fadd_rn(acc, cur[k]);
                        cur[k] = acc;
                    }
                    tm_st32(tout + g * 32, cur);
                };
#pragma unroll
                for (int g = 0; g < kT3Unit / 32; g += 2) {
                    group(g, va, vb);
                    group(g + 1, vb, va);
                }
                tm_wait_st();
                tm_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(tempty0 + 8 * ts);   // the IN slot has been read (the last wait::ld is behind us)
                    mbar_arrive(ofull0 + 8 * ts);
                }
            }
            for (int u = (nunits > kT3TRing ? nunits - kT3TRing : 0); u < nunits; u++) {   // drain: phases in step for the next group
                const int ts = u % kT3TRing;
                mbar_spin(oempty0 + 8 * ts, (ph2 >> ts) & 1u);
                ph2 ^= 1u << ts;
            }
        } else if (is_stager) {
            // ---- stager: shared-memory ring -> Tensor Memory IN ring, lane = (chain lane / 2)
            const int ch = lane >> kT3LaneShift;
            for (int u = 0; u < nunits; u++) {
                const int r = u % kT3Ring, ts = u % kT3TRing;
                mbar_wait(full0 + 8 * r, (ph >> r) & 1u);
                ph ^= 1u << r;
                if (u >= kT3TRing) {   // the adder has read this IN slot's previous unit
                    mbar_wait(tempty0 + 8 * ts, (ph2 >> ts) & 1u);
                    ph2 ^= 1u << ts;
                    tm_fence_after();
                }
#pragma unroll
                for (int g = 0; g < kT3Unit / 32; g++) {
                    float v[32];
                    const float4* mv = reinterpret_cast<const float4*>(&m_s[r][ch][g * 32]);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float4 q = mv[k];
                        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
                    }
                    tm_st32(tin0 + ts * kT3Unit + g * 32, v);
                }
                tm_wait_st();
                tm_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(empty0 + 8 * r);
                    mbar_arrive(tfull0 + 8 * ts);
                }
            }
            for (int u = (nunits > kT3TRing ? nunits - kT3TRing : 0); u < nunits; u++) {
                const int ts = u % kT3TRing;
                mbar_wait(tempty0 + 8 * ts, (ph2 >> ts) & 1u);
                ph2 ^= 1u << ts;
            }
        } else if (is_writer) {
            // ---- writer: Tensor Memory OUT ring -> the chains' scratch rows (whole vectors: the rows are padded to whole units)
            const int ch = lane >> kT3LaneShift;
            const bool active = (lane & ((1 << kT3LaneShift) - 1)) == 0 && slot0 + ch < nslots;
            float* out = r900_row(scratch, slot0 + (active ? ch : 0), span);   // out[i] = csum[i]
            if (active) out[0] = 0.0f;
            for (int u = 0; u < nunits; u++) {
                const int ts = u % kT3TRing;
                mbar_wait(ofull0 + 8 * ts, (ph >> ts) & 1u);
                ph ^= 1u << ts;
                tm_fence_after();
#pragma unroll
                for (int g = 0; g < kT3Unit / 32; g++) {
                    float v[32];
                    tm_ld32(tout0 + ts * kT3Unit + g * 32, v);
                    tm_wait_ld32(v);
                    if (active) {
                        float4* o4 = reinterpret_cast<float4*>(out + u * kT3Unit + g * 32 + 1);   // csum[i0 + 1 ..]: 16-byte aligned
#pragma unroll
                        for (int k = 0; k < 8; k++) o4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    }
                }
                tm_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(oempty0 + 8 * ts);
            }
        } else if (prod < kT3Producers) {
            // ---- producers: as in r900_chain2_kernel (asynchronous copies, computed magnitudes, STS into the ring)
            const int ch = prod * kT3ChainsPerWarp + lane / kT3Lanes, off = (lane % kT3Lanes) * 8;
            const unsigned int slot = slot0 + ch;
            const bool valid = slot < nslots;
            const long long first = ((long long)(valid ? slot_block[slot] : 0) + 1) * cfg.BS - cfg.BUF;   // sample of the parser's signal[0]
            const uint32_t ring0 = smem_u32(&raw_s[prod][0][0][lane]);
            auto fetch = [&](int u) {
#pragma unroll
                for (int l = 0; l < kT3Loads; l++) {
                    const int i = u * kT3Unit + l * (8 * kT3Lanes) + off;
                    cp_async16(ring0 + (uint32_t)((u % kT3Ahead) * kT3Loads + l) * 32u * 16u,
                               raw8_addr(iq, hist, hist_samples, first + i, valid && u < nunits && i < span - 1));
                }
                cp_async_commit();
            };
            for (int a = 0; a < kT3Ahead; a++) fetch(a);
            for (int u = 0; u < nunits; u++) {
                const int r = u % kT3Ring;
                cp_async_wait<kT3Ahead - 1>();
                float m[kT3Loads][8];
#pragma unroll
                for (int l = 0; l < kT3Loads; l++) {
                    const uint4 rv = raw_s[prod][u % kT3Ahead][l][lane];
                    const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float i0 = lut_of_magic(__uint_as_float(__byte_perm(w[k], 0x47000000u, 0x7604)));
                        const float q0 = lut_of_magic(__uint_as_float(__byte_perm(w[k], 0x47000000u, 0x7614)));
                        const float i1 = lut_of_magic(__uint_as_float(__byte_perm(w[k], 0x47000000u, 0x7624)));
                        const float q1 = lut_of_magic(__uint_as_float(__byte_perm(w[k], 0x47000000u, 0x7634)));
                        m[l][2 * k] = __fadd_rn(i0, q0);
                        m[l][2 * k + 1] = __fadd_rn(i1, q1);
                    }
                    const int i0s = u * kT3Unit + l * (8 * kT3Lanes) + off;
                    const long long j0 = first + i0s;
                    const bool whole = valid && i0s + 8 <= span - 1 && (j0 >= 0 || -j0 <= (long long)hist_valid);
                    if (!whole) {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const long long j = j0 + k;
                            const bool ok = valid && i0s + k < span - 1 && (j >= 0 || -j <= (long long)hist_valid);
                            if (!ok) m[l][k] = 0.0f;
                        }
                    }
                }
                fetch(u + kT3Ahead);
                if (u >= kT3Ring) {
                    mbar_wait(empty0 + 8 * r, (ph >> r) & 1u);
                    ph ^= 1u << r;
                }
#pragma unroll
                for (int l = 0; l < kT3Loads; l++) {
                    float4* d4 = reinterpret_cast<float4*>(&m_s[r][ch][l * (8 * kT3Lanes) + off]);
                    d4[0] = make_float4(m[l][0], m[l][1], m[l][2], m[l][3]);
                    d4[1] = make_float4(m[l][4], m[l][5], m[l][6], m[l][7]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * r);
            }
            cp_async_wait<0>();
            for (int u = (nunits > kT3Ring ? nunits - kT3Ring : 0); u < nunits; u++) {
                const int r = u % kT3Ring;
                mbar_wait(empty0 + 8 * r, (ph >> r) & 1u);
                ph ^= 1u << r;
            }
        }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "r"(kT3Cols) : "memory");
}

#undef ERT_R32

}  // namespace ert
