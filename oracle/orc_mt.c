/*
 * orc_mt.c — pthread drivers over the oracle cascades (TEST INFRASTRUCTURE).
 * Only used for the `cpu_baseline.kind == "port"` fallback of bench.py when
 * oracle/_ref/ is not available; channels are split evenly across threads
 * (the cascade functions are re-entrant across distinct state pointers,
 * SURVEY.md §8b "Threading").
 */
#include <pthread.h>
#include <time.h>
#include "dspi_oracle.h"

typedef struct { int kind; void *bq; void *s; uint32_t c0, c1, T, nbands, packet; } job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    uint32_t C = j->c1 - j->c0;
    if (j->kind == 2)
        orc_q28_eq_many((orc_biquad_q28 *)j->bq + (size_t)j->c0 * ORC_MAX_BANDS, (int32_t *)j->s + (size_t)j->c0 * j->T, C, j->T, j->nbands, j->packet);
    else if (j->kind == 1)
        orc_f32s_eq_many((orc_biquad_f32 *)j->bq + (size_t)j->c0 * ORC_MAX_BANDS, (float *)j->s + (size_t)j->c0 * j->T, C, j->T, j->nbands, j->packet);
    else
        orc_f32f_eq_many((orc_biquad_f32 *)j->bq + (size_t)j->c0 * ORC_MAX_BANDS, (float *)j->s + (size_t)j->c0 * j->T, C, j->T, j->nbands, j->packet);
    return NULL;
}

/* kind: 0 f32 fused, 1 f32 strict, 2 q28.  Returns elapsed seconds. */
double orc_eq_many_mt(int kind, void *bq, void *samples, uint32_t C, uint32_t T, uint32_t nbands, uint32_t packet, uint32_t nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    job_t jobs[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t i = 0; i < nthreads; i++) {
        jobs[i] = (job_t){ kind, bq, samples, (uint32_t)((uint64_t)C * i / nthreads), (uint32_t)((uint64_t)C * (i + 1) / nthreads), T, nbands, packet };
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    for (uint32_t i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
