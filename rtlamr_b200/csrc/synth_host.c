/* synth_host.c -- gcc-built host twin of the CUDA generator (bench/test tooling). */
#include "../../include/ertgpu_synth.h"

void ertsynth_host_fill(uint8_t *out, int64_t first_sample, int64_t nsamples, uint64_t seed,
                        const ertgpu_synth_packet *pk, int64_t npk) {
    synth_reference_fill(out, first_sample, nsamples, seed, pk, npk);
}
