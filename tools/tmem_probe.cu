// tmem_probe.cu -- can Tensor Memory serve as a per-thread ring scratchpad next to a register-bound kernel?
// Measures tcgen05.ld / tcgen05.st (.32x32b.x8) throughput and round-trip correctness with W warps per CTA,
// one CTA per SM, no MMA anywhere.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld8(uint32_t a, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a));
}
__device__ __forceinline__ void tm_st8(uint32_t a, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: ring walk (ld slot group, wait, add, st other slot group), as the demod kernel would do it
// MODE 1: ld only    MODE 2: st only    MODE 3: correctness: st then ld of the same columns without wait::st
template <int MODE>
__global__ void probe(int iters, int cols_per_warp, unsigned long long* cyc, uint32_t* bad, float* sink) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    // this warp's lane quarter and its share of the 512 columns
    const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * cols_per_warp);
    uint32_t r[8], s[8];
    for (int i = 0; i < 8; i++) { r[i] = 0; s[i] = threadIdx.x * 1000 + i; }
    // initialise the ring
    for (int c = 0; c < cols_per_warp; c += 8) {
        for (int i = 0; i < 8; i++) s[i] = (threadIdx.x << 16) | (uint32_t)(c + i);
        tm_st8(my + c, s);
    }
    tm_wait_st();
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    uint32_t nbad = 0;
    int c = 0;
    for (int it = 0; it < iters; it++) {
        int c2 = c + 16; if (c2 >= cols_per_warp) c2 -= cols_per_warp;
        if (MODE == 0) {
            tm_ld8(my + c, r);
            tm_wait_ld();
            for (int i = 0; i < 8; i++) { acc += __uint_as_float(r[i] & 0x3FFFFFFF); s[i] = r[i] ^ 0x5u; }
            tm_st8(my + c2, s);
        } else if (MODE == 1) {
            tm_ld8(my + c, r);
            tm_wait_ld();
            for (int i = 0; i < 8; i++) acc += __uint_as_float(r[i] & 0x3FFFFFFF);
        } else if (MODE == 2) {
            for (int i = 0; i < 8; i++) s[i] += it;
            tm_st8(my + c, s);
        } else {
            for (int i = 0; i < 8; i++) s[i] = (uint32_t)it * 977u + (threadIdx.x << 8) + i + blockIdx.x;
            tm_st8(my + c, s);
            tm_ld8(my + c, r);          // no wait::st in between: is the pair ordered?
            tm_wait_ld();
            for (int i = 0; i < 8; i++) nbad += (r[i] != s[i]);
        }
        c += 8; if (c >= cols_per_warp) c = 0;
    }
    tm_wait_st();
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 32 + warp] = (unsigned long long)(t1 - t0);
    if (nbad) atomicAdd(bad, nbad);
    if (acc == 123.456f) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long* cyc; uint32_t* bad; float* sink;
    cudaMalloc(&cyc, sms * 32 * 8); cudaMalloc(&bad, 4); cudaMalloc(&sink, 4);
    const int iters = 20000;
    for (int mode = 0; mode < 4; mode++)
        for (int W : {4, 8, 12, 16}) {
            const int cpw = 512 / ((W + 3) / 4) / 8 * 8;
            cudaMemset(bad, 0, 4);
            cudaMemset(cyc, 0, sms * 32 * 8);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            switch (mode) {
                case 0: probe<0><<<sms, W * 32>>>(iters, cpw, cyc, bad, sink); break;
                case 1: probe<1><<<sms, W * 32>>>(iters, cpw, cyc, bad, sink); break;
                case 2: probe<2><<<sms, W * 32>>>(iters, cpw, cyc, bad, sink); break;
                default: probe<3><<<sms, W * 32>>>(iters, cpw, cyc, bad, sink); break;
            }
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            unsigned long long h[32]; uint32_t hb = 0;
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (int i = 0; i < W; i++) if (h[i] > mx) mx = h[i];
            // bytes per SM per cycle: W warps x iters x 8 regs x 128 B (x2 for mode 0: ld + st)
            const double bytes = (double)W * iters * 8 * 128 * (mode == 0 ? 2 : 1) * (mode == 3 ? 2 : 1);
            printf("mode %d W=%2d cols/warp=%3d: %s  %.3f ms  %llu cyc  %.1f B/cyc/SM  %.2f cyc per warp-iteration(SM-wide)  bad=%u\n", mode, W, cpw,
                   cudaGetErrorString(e), ms, mx, bytes / (double)mx, (double)mx / (iters * (double)W), hb);
        }
    return 0;
}
