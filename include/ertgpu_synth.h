/*
 * ertgpu_synth.h -- the synthetic IQ stream used by bench.py and the parity
 * tests, as one header shared by the CUDA generator (rtlamr_b200/csrc/synth.cu)
 * and by host code (gcc), so both produce bit-identical bytes for any sharding.
 *
 * Not part of the reference API (rtlamr has no signal generator); the layout
 * follows SURVEY.md section 8d "Synthetic input":
 *   noise:   counter-based hash keyed by (seed, global sample index); I and Q
 *            are each 127 + round(s/32) where s is a sum of four uniform bytes
 *            minus 510 (Irwin-Hall, sigma ~ 4.6 LSB) -- integer-only so host and
 *            device agree exactly;
 *   packets: on "high" chips the carrier (amp_i, amp_q) is added to (I, Q);
 *            "low" chips are noise only (OOK).  Values clamp to [0,255].
 */
#ifndef ERTGPU_SYNTH_H
#define ERTGPU_SYNTH_H

#include <stdint.h>

#include "ertgpu.h"

#if defined(__CUDACC__)
#define ERTSYNTH_HD __host__ __device__ __forceinline__
#else
#define ERTSYNTH_HD static inline
#endif

ERTSYNTH_HD uint64_t ertsynth_mix(uint64_t x) {
    /* splitmix64 finaliser */
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

/* noise-only I,Q (before clamping; always inside [96,159]) */
ERTSYNTH_HD void ertsynth_noise(uint64_t seed, int64_t n, int32_t *i, int32_t *q) {
    uint64_t h = ertsynth_mix(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(n + 1));
    int32_t si = (int32_t)(h & 0xFF) + (int32_t)((h >> 8) & 0xFF) + (int32_t)((h >> 16) & 0xFF) +
                 (int32_t)((h >> 24) & 0xFF) - 510;
    int32_t sq = (int32_t)((h >> 32) & 0xFF) + (int32_t)((h >> 40) & 0xFF) +
                 (int32_t)((h >> 48) & 0xFF) + (int32_t)((h >> 56) & 0xFF) - 510;
    /* floor((s+16)/32) with a shift of a non-negative value: s+16+1024 >= 0 */
    *i = 127 + ((si + 16 + 1024) >> 5) - 32;
    *q = 127 + ((sq + 16 + 1024) >> 5) - 32;
}

ERTSYNTH_HD int32_t ertsynth_clamp(int32_t v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* index of the last packet with start_sample <= n, or -1 */
ERTSYNTH_HD int64_t ertsynth_find(const ertgpu_synth_packet *pk, int64_t npk, int64_t n) {
    int64_t lo = 0, hi = npk; /* first packet with start > n */
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (pk[mid].start_sample <= n) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

/* full sample: noise + packet carrier */
ERTSYNTH_HD void ertsynth_sample(uint64_t seed, int64_t n, const ertgpu_synth_packet *pk,
                                 int64_t cand, uint8_t *out_i, uint8_t *out_q) {
    int32_t i, q;
    ertsynth_noise(seed, n, &i, &q);
    if (cand >= 0) {
        const ertgpu_synth_packet *p = &pk[cand];
        int64_t off = n - p->start_sample;
        if (off >= 0 && off < (int64_t)p->n_chips * p->chip_length) {
            int32_t chip = (int32_t)(off / p->chip_length);
            if ((p->chips[chip >> 3] >> (7 - (chip & 7))) & 1) {
                i += p->amp_i;
                q += p->amp_q;
            }
        }
    }
    *out_i = (uint8_t)ertsynth_clamp(i);
    *out_q = (uint8_t)ertsynth_clamp(q);
}

#if !defined(__CUDACC__)
/* host reference generator: bit-identical to ertgpu_synth_fill */
static inline void synth_reference_fill(uint8_t *out, int64_t first_sample, int64_t nsamples,
                                        uint64_t seed, const ertgpu_synth_packet *pk, int64_t npk) {
    for (int64_t k = 0; k < nsamples; k++) {
        int64_t n = first_sample + k;
        int64_t c = ertsynth_find(pk, npk, n);
        ertsynth_sample(seed, n, pk, c, &out[2 * k], &out[2 * k + 1]);
    }
}
#endif

#endif
