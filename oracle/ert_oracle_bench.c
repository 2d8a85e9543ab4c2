/*
 * ert_oracle_bench.c -- timed CPU baseline driver around the oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see ert_oracle.h): used by bench.py's cpu_baseline
 * leg and by `bench.py --impl reference`, nowhere else.
 *
 * The reference runs its DSP on ONE goroutine (main.go:207-235: one Decode per
 * BlockSize2 bytes, drain the message channel, next block).  The "whole host"
 * figure is T such decoders, each an independent rtlamr process in effect,
 * over T contiguous block-aligned shards of ONE stream: thread t owns blocks
 * [t*n/T, (t+1)*n/T) and calls Decode once per block exactly like main.go:235.
 * No halo is fed (a throughput measurement: the candidates next to a seam are
 * not compared with anything here); everything a Decode call allocates in the
 * reference per block (candidate slices, message structs) is pre-allocated
 * outside the timed region, which favours this baseline over the Go original.
 * Threads are created, pinned round-robin over the CPUs the process may run on
 * and parked on a flag before the clock starts; the clock stops when the
 * last thread finishes.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ert_oracle.h"

typedef struct {
    ert_oracle *o;
    const uint8_t *iq;
    int64_t first_block, nblocks;
    int32_t repeats;
    int cpu; /* -1: not pinned */
    volatile int *go; /* 0 = parked, 1 = run, 2 = give up */
    volatile int ready;
    ert_oracle_cand *cands;
    ert_oracle_msg *msgs;
    int32_t cand_cap, msg_cap;
    int64_t ncands, nmsgs;
    double t_end;
} bench_thread;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *bench_main(void *arg) {
    bench_thread *t = (bench_thread *)arg;
    if (t->cpu >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(t->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    const size_t bs2 = (size_t)ert_oracle_config(t->o)->block_size2;
    t->ready = 1;
    while (*t->go == 0) sched_yield();
    if (*t->go != 1) return NULL;
    for (int32_t r = 0; r < t->repeats; r++) {
        const uint8_t *p = t->iq + (size_t)t->first_block * bs2;
        for (int64_t b = 0; b < t->nblocks; b++, p += bs2) {
            int32_t nc = 0, nm = 0; /* main.go:235-277: the block's messages are consumed before the next Decode */
            ert_oracle_decode(t->o, p, t->cands, t->cand_cap, &nc, t->msgs, t->msg_cap, &nm);
            t->ncands += nc;
            t->nmsgs += nm;
        }
    }
    t->t_end = now_s();
    return NULL;
}

/* Returns the wall seconds from the common start to the last thread's finish (< 0 on failure).
 * ncands/nmsgs: totals over all threads and repeats. */
double ert_oracle_bench_threads(const int32_t *protos, int32_t nprotos, int32_t chip_length,
                                int32_t search_mode, const uint8_t *iq, int64_t nblocks,
                                int32_t nthreads, int32_t repeats, int32_t pin,
                                int64_t *ncands, int64_t *nmsgs) {
    if (nthreads < 1 || nblocks < 0 || repeats < 1) return -1.0;
    if (nthreads > nblocks && nblocks > 0) nthreads = (int32_t)nblocks;
    bench_thread *th = (bench_thread *)calloc((size_t)nthreads, sizeof(*th));
    pthread_t *ids = (pthread_t *)calloc((size_t)nthreads, sizeof(*ids));
    volatile int go = 0;
    if (!th || !ids) return -1.0;

    /* CPUs this process may use, in order */
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    int ncpu = 0, cpus[CPU_SETSIZE];
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;

    int ok = 1;
    for (int32_t i = 0; i < nthreads; i++) {
        bench_thread *t = &th[i];
        t->o = ert_oracle_new(protos, nprotos, chip_length, search_mode);
        if (!t->o) { ok = 0; break; }
        t->iq = iq;
        t->first_block = nblocks * i / nthreads;
        t->nblocks = nblocks * (i + 1) / nthreads - t->first_block;
        t->repeats = repeats;
        t->cpu = (pin && ncpu > 0) ? cpus[i % ncpu] : -1;
        t->go = &go;
        t->cand_cap = 4 * ert_oracle_config(t->o)->block_size; /* Search returns at most BlockSize starts per preamble */
        t->msg_cap = 256;
        t->cands = (ert_oracle_cand *)malloc((size_t)t->cand_cap * sizeof(*t->cands));
        t->msgs = (ert_oracle_msg *)malloc((size_t)t->msg_cap * sizeof(*t->msgs));
        if (!t->cands || !t->msgs) { ok = 0; break; }
    }
    double dt = -1.0;
    if (ok) {
        int32_t started = 0;
        for (; started < nthreads; started++)
            if (pthread_create(&ids[started], NULL, bench_main, &th[started]) != 0) break;
        if (started == nthreads) {
            for (int32_t i = 0; i < nthreads; i++)
                while (!th[i].ready) sched_yield(); /* every thread is pinned and parked */
            const double t0 = now_s();
            go = 1;
            double t1 = t0;
            for (int32_t i = 0; i < nthreads; i++) {
                pthread_join(ids[i], NULL);
                if (th[i].t_end > t1) t1 = th[i].t_end;
            }
            dt = t1 - t0;
        } else {
            go = 2; /* could not start every thread: release the parked ones */
            for (int32_t i = 0; i < started; i++) pthread_join(ids[i], NULL);
        }
    }
    int64_t nc = 0, nm = 0;
    for (int32_t i = 0; i < nthreads; i++) {
        nc += th[i].ncands;
        nm += th[i].nmsgs;
        free(th[i].cands);
        free(th[i].msgs);
        ert_oracle_free(th[i].o);
    }
    if (ncands) *ncands = nc;
    if (nmsgs) *nmsgs = nm;
    free(th);
    free(ids);
    return dt;
}

/* CPUs the calling process may run on (len(os.sched_getaffinity(0))) */
int32_t ert_oracle_host_cpus(void) {
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return 1;
    int n = CPU_COUNT(&allowed);
    return n > 0 ? n : 1;
}
