// pipe_probe.cu -- issue cost of the demod kernel's instruction classes on sm_100a, alone and mixed:
// does an ALU-pipe instruction (PRMT/SHF/LOP3) overlap with an FMA-pipe one (FADD/FADD2/FFMA/IDP), or do they share
// the dispatch port?  One CTA per SM, W warps per scheduler, 8 independent dependency chains per class.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MIX>
__global__ void probe(int iters, unsigned long long* cyc, uint32_t* sink, uint32_t seed) {
    uint32_t a[8], b[8];
    float f[8], g[8];
    uint64_t p[8];
    for (int i = 0; i < 8; i++) {
        a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed ^ (i * 77u);
        f[i] = (float)(i + 1) * 0.001f; g[i] = 1.0f + i;
        p[i] = ((uint64_t)__float_as_uint(f[i]) << 32) | __float_as_uint(g[i]);
    }
    const uint32_t k1 = seed | 0x3210u;
    const float big = 1.152921504606847e18f, one = 1.0f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#define PRMT(i) asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(a[i]) : "r"(k1));
#define SHF(i) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(b[i]) : "r"(a[i]));
#define FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
#define FADD2(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
#define FFMASAT(i) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(big), "f"(one));
#define FFMAI(i) asm volatile("fma.rn.f32 %0, %0, 0f40000000, %1;" : "+f"(g[i]) : "f"(f[i]));
#define IDP(i) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(0x80u), "r"(k1));
#define IMAD(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(b[i]));
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(k1));
        if (MIX == 0) { REP8(PRMT) REP8(PRMT) }
        if (MIX == 1) { REP8(FADD) REP8(FADD) }
        if (MIX == 2) { REP8(FADD2) REP8(FADD2) }
        if (MIX == 3) { REP8(SHF) REP8(SHF) }
        if (MIX == 4) { REP8(IDP) REP8(IDP) }
        if (MIX == 5) { REP8(IMAD) REP8(IMAD) }
        if (MIX == 6) { REP8(FFMASAT) REP8(FFMASAT) }
        if (MIX == 7) { REP8(FFMAI) REP8(FFMAI) }
        // mixes: 16 instructions of two classes interleaved
#define PF(i) PRMT(i) FADD(i)
#define PF2(i) PRMT(i) FADD2(i)
#define PI(i) PRMT(i) IDP(i)
#define PL(i) PRMT(i) LOP(i)
#define IF(i) IDP(i) FADD(i)
#define IF2(i) IDP(i) FADD2(i)
#define FF2(i) FADD(i) FADD2(i)
#define SF(i) FFMASAT(i) FFMAI(i)
        if (MIX == 8) { REP8(PF) }
        if (MIX == 9) { REP8(PF2) }
        if (MIX == 10) { REP8(PI) }
        if (MIX == 11) { REP8(PL) }
        if (MIX == 12) { REP8(IF) }
        if (MIX == 13) { REP8(IF2) }
        if (MIX == 14) { REP8(FF2) }
        if (MIX == 15) { REP8(SF) }
        // the demod mix per 2 samples: 4 PRMT, 2 SHF, 2 FADD, 3 FADD2 (no loads)
        if (MIX == 16) { PRMT(0) PRMT(1) PRMT(2) PRMT(3) FADD2(0) FADD(0) FADD(1) FADD2(1) FADD2(2) SHF(0) SHF(1)
                         PRMT(4) PRMT(5) PRMT(6) PRMT(7) FADD2(3) FADD(2) FADD(3) FADD2(4) FADD2(5) SHF(2) SHF(3) }
        // same with IDP addresses and FFMA sign accumulation
        if (MIX == 17) { IDP(0) IDP(1) IDP(2) IDP(3) FADD2(0) FADD(0) FADD(1) FADD2(1) FADD2(2) FFMASAT(4) FFMAI(4) FFMASAT(5) FFMAI(5)
                         IDP(4) IDP(5) IDP(6) IDP(7) FADD2(3) FADD(2) FADD(3) FADD2(4) FADD2(5) FFMASAT(6) FFMAI(6) FFMASAT(7) FFMAI(7) }
    }
    const long long t1 = clock64();
    uint32_t x = 0;
    for (int i = 0; i < 8; i++) x ^= a[i] ^ b[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]) ^ (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32);
    if (x == 0x9E3779B9u) sink[0] = x;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = (unsigned long long)(t1 - t0);
}

static const int kN[18] = {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 22, 26};
static const char* kName[18] = {"PRMT", "FADD", "FADD2", "SHF", "IDP4A", "IMAD", "FFMA.SAT", "FFMA imm", "PRMT+FADD", "PRMT+FADD2", "PRMT+IDP", "PRMT+LOP3",
                                "IDP+FADD", "IDP+FADD2", "FADD+FADD2", "FFMA.SAT+FFMAimm", "demod mix (4 PRMT 2 SHF 2 FADD 3 FADD2) x2", "IDP/FFMA mix x2"};

template <int MIX>
void run(unsigned long long* cyc, uint32_t* sink) {
    const int iters = 4000;
    for (int wps : {1, 2, 3, 4}) {   // warps per scheduler
        probe<MIX><<<148, wps * 128>>>(iters, cyc, sink, 12345u);
        cudaDeviceSynchronize();
        unsigned long long h[32];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (int i = 0; i < wps * 4; i++) if (h[i] > mx) mx = h[i];
        // cycles per instruction per scheduler
        printf("%-46s %d warps/sched: %.2f cyc per warp-instruction (per scheduler)\n", kName[MIX], wps, (double)mx / ((double)iters * kN[MIX] * wps));
    }
}

int main() {
    unsigned long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 32 * 8); cudaMalloc(&sink, 4);
    run<0>(cyc, sink); run<1>(cyc, sink); run<2>(cyc, sink); run<3>(cyc, sink); run<4>(cyc, sink); run<5>(cyc, sink); run<6>(cyc, sink); run<7>(cyc, sink);
    run<8>(cyc, sink); run<9>(cyc, sink); run<10>(cyc, sink); run<11>(cyc, sink); run<12>(cyc, sink); run<13>(cyc, sink); run<14>(cyc, sink); run<15>(cyc, sink);
    run<16>(cyc, sink); run<17>(cyc, sink);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
