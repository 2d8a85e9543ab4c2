/* synth_host.c -- gcc-built host twin of the CUDA generator (bench/test tooling). */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>

#include "../../include/ertgpu_synth.h"

void ertsynth_host_fill(uint8_t *out, int64_t first_sample, int64_t nsamples, uint64_t seed,
                        const ertgpu_synth_packet *pk, int64_t npk) {
    synth_reference_fill(out, first_sample, nsamples, seed, pk, npk);
}

/* the same bytes from `nthreads` threads (the generator is counter based: any range, any order) */
typedef struct {
    uint8_t *out;
    int64_t first_sample, nsamples;
    uint64_t seed;
    const ertgpu_synth_packet *pk;
    int64_t npk;
} fill_job;

static void *fill_main(void *arg) {
    fill_job *j = (fill_job *)arg;
    synth_reference_fill(j->out, j->first_sample, j->nsamples, j->seed, j->pk, j->npk);
    return NULL;
}

void ertsynth_host_fill_mt(uint8_t *out, int64_t first_sample, int64_t nsamples, uint64_t seed,
                           const ertgpu_synth_packet *pk, int64_t npk, int32_t nthreads) {
    if (nthreads < 2 || nsamples < (1 << 16)) {
        synth_reference_fill(out, first_sample, nsamples, seed, pk, npk);
        return;
    }
    fill_job *jobs = (fill_job *)calloc((size_t)nthreads, sizeof(*jobs));
    pthread_t *ids = (pthread_t *)calloc((size_t)nthreads, sizeof(*ids));
    int32_t started = 0;
    for (int32_t t = 0; jobs && ids && t < nthreads; t++) {
        const int64_t a = nsamples * t / nthreads, b = nsamples * (t + 1) / nthreads;
        jobs[t] = (fill_job){out + 2 * a, first_sample + a, b - a, seed, pk, npk};
        if (pthread_create(&ids[t], NULL, fill_main, &jobs[t]) != 0) {
            fill_main(&jobs[t]); /* no thread: do the slice here */
            ids[t] = 0;
        } else {
            started++;
        }
    }
    if (!jobs || !ids) synth_reference_fill(out, first_sample, nsamples, seed, pk, npk);
    else
        for (int32_t t = 0; t < nthreads; t++)
            if (ids[t]) pthread_join(ids[t], NULL);
    (void)started;
    free(jobs);
    free(ids);
}
