// synth.cuh -- device generator of the synthetic IQ stream (bench/test tooling).
// Bit-identical to synth_reference_fill() in include/ertgpu_synth.h.
#pragma once

#include "../../include/ertgpu_synth.h"

namespace ert {

// each thread produces 8 consecutive samples (16 bytes, one vector store)
__global__ void synth_kernel(uint8_t* __restrict__ out, long long first_sample, long long nsamples,
                             unsigned long long seed, const ertgpu_synth_packet* __restrict__ pk,
                             long long npk) {
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; k0 < nsamples; k0 += stride) {
        long long cand = ertsynth_find(pk, npk, first_sample + k0);
        uint8_t v[16];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const long long n = first_sample + k0 + u;
            // the next packet may start inside this group of 8
            if (cand + 1 < npk && pk[cand + 1].start_sample <= n) cand++;
            ertsynth_sample(seed, n, pk, cand, &v[2 * u], &v[2 * u + 1]);
        }
        if (k0 + 8 <= nsamples && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            uint4 w;
            memcpy(&w, v, 16);
            reinterpret_cast<uint4*>(out)[k0 >> 3] = w;
        } else {
            for (int u = 0; u < 8 && k0 + u < nsamples; u++) {
                out[2 * (k0 + u)] = v[2 * u];
                out[2 * (k0 + u) + 1] = v[2 * u + 1];
            }
        }
    }
}

}  // namespace ert
