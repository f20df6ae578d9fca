/*
 * ref_shim.c — thin C entry points around the UNMODIFIED reference sources
 * (TEST INFRASTRUCTURE).  Compiled together with
 *   /root/reference/firmware/DSPi/{dsp_pipeline,crossfeed,leveller,loudness}.c
 * straight from where they lie (oracle/Makefile target `ref`), into
 *   oracle/_ref/libdspi_ref_f32_strict.so   -DPICO_RP2350=1 -ffp-contract=off
 *   oracle/_ref/libdspi_ref_f32_fused.so    -DPICO_RP2350=1 -mfma -ffp-contract=fast
 *   oracle/_ref/libdspi_ref_q28.so          -DPICO_RP2350=0 -fwrapv
 * No reference source is copied into this repository; this file only calls
 * the reference's public functions and reports its struct geometry so the
 * tests can pin the restatement (oracle/dspi_oracle.h) against it.
 */
#include <stddef.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#if defined(__x86_64__)
#include <xmmintrin.h>
#endif
#include "config.h"
#include "dsp_pipeline.h"
#include "crossfeed.h"
#include "leveller.h"
#include "loudness.h"

#if PICO_RP2350
typedef float sample_t;
#else
typedef int32_t sample_t;
#endif

static inline unsigned ftz_enter(void)
{
#if defined(__x86_64__)
    unsigned old = _mm_getcsr();
    _mm_setcsr(old | 0x8040u);      /* FTZ|DAZ == firmware FPSCR FZ+DN, main.c:593-600 */
    return old;
#else
    return 0;
#endif
}
static inline void ftz_leave(unsigned old)
{
#if defined(__x86_64__)
    _mm_setcsr(old);
#else
    (void)old;
#endif
}

#if !PICO_RP2350
/* The RP2040 cascade lives in dsp_process_rp2040.S (Thumb-1) and cannot be
 * assembled for the host.  This stand-in follows the assembly's block routine
 * (dsp_process_rp2040.S:225-394) but multiplies with the REFERENCE's own
 * fast_mul_q28() (dsp_pipeline.c:47-58, compiled unmodified): each inlined
 * multiply in the assembly (e.g. :273-283) is instruction-for-instruction that
 * function.  It is an independent second statement of the routine, used to
 * cross-check oracle/orc_q28.c. */
void dsp_process_channel_block(Biquad *restrict biquads, int32_t *restrict samples, uint32_t count, uint8_t channel)
{
    uint8_t nb = channel_band_counts[channel];
    for (uint8_t band = 0; band < nb; band++) {
        Biquad *bq = &biquads[band];
        if (bq->bypass) continue;
        int32_t s1 = bq->s1, s2 = bq->s2;
        for (uint32_t i = 0; i < count; i++) {
            int32_t x = samples[i];
            int32_t y = fast_mul_q28(bq->b0, x) + s1;
            int32_t t1 = fast_mul_q28(bq->b1, x);
            int32_t t3 = fast_mul_q28(bq->b2, x);
            s1 = (t1 - fast_mul_q28(bq->a1, y)) + s2;
            s2 = t3 - fast_mul_q28(bq->a2, y);
            samples[i] = y;
        }
        bq->s1 = s1;
        bq->s2 = s2;
    }
}
int32_t dsp_process_channel(Biquad *restrict biquads, int32_t input, uint8_t channel)
{
    dsp_process_channel_block(biquads, &input, 1, channel);
    return input;
}
int32_t ref_mul_q28(int32_t a, int32_t b) { return fast_mul_q28(a, b); }
int32_t ref_mul_q15(int32_t a, int32_t b) { return fast_mul_q15(a, b); }
#endif

int ref_is_float(void) { return PICO_RP2350 ? 1 : 0; }

size_t ref_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(Biquad);
    case 1: return sizeof(EqParamPacket);
    case 2: return sizeof(MatrixCrosspoint);
    case 3: return sizeof(OutputChannel);
    case 4: return sizeof(MatrixMixer);
    case 5: return sizeof(CrossfeedState);
    case 6: return sizeof(LevellerCoeffs);
    case 7: return sizeof(LevellerState);
    case 8: return sizeof(LoudnessCoeffs);
    case 9: return sizeof(CrossfeedConfig);
    case 10: return sizeof(LevellerConfig);
    case 11: return MAX_BANDS;
    case 12: return NUM_CHANNELS;
    case 13: return NUM_OUTPUT_CHANNELS;
    case 14: return MAX_DELAY_SAMPLES;
    case 15: return sizeof(SystemStatusPacket);
    default: return 0;
    }
}
size_t ref_offsetof(int which)
{
    switch (which) {
    case 0: return offsetof(Biquad, s1);
    case 1: return offsetof(Biquad, bypass);
#if PICO_RP2350
    case 2: return offsetof(Biquad, sva1);
    case 3: return offsetof(Biquad, svm0);
    case 4: return offsetof(Biquad, svic1eq);
    case 5: return offsetof(Biquad, svf_type);
    case 6: return offsetof(Biquad, use_svf);
#endif
    case 7: return offsetof(LevellerState, la_write_idx);
    case 8: return offsetof(LevellerConfig, gate_threshold_db);
    case 9: return offsetof(CrossfeedConfig, custom_feed_db);
    case 10: return offsetof(MatrixMixer, outputs);
    default: return (size_t)-1;
    }
}

void ref_set_nbands(uint32_t nbands)
{
    for (int c = 0; c < NUM_CHANNELS; c++) channel_band_counts[c] = (uint8_t)nbands;
}

void ref_eq_coeffs(EqParamPacket *p, Biquad *bq, float fs) { dsp_compute_coefficients(p, bq, fs); }

void ref_eq_block(Biquad *bq, sample_t *samples, uint32_t count)
{
    unsigned csr = ftz_enter();
    dsp_process_channel_block(bq, samples, count, 0);
    ftz_leave(csr);
}

/* bq[C][MAX_BANDS], samples[C][T]; firmware-style packets */
void ref_eq_many(Biquad *bq, sample_t *samples, uint32_t C, uint32_t T, uint32_t packet)
{
    unsigned csr = ftz_enter();
    if (packet == 0) packet = T;
    for (uint32_t c = 0; c < C; c++) {
        Biquad *b = bq + (size_t)c * MAX_BANDS;
        sample_t *s = samples + (size_t)c * T;
        for (uint32_t t0 = 0; t0 < T; t0 += packet) {
            uint32_t n = (T - t0 < packet) ? (T - t0) : packet;
            dsp_process_channel_block(b, s + t0, n, 0);
        }
    }
    ftz_leave(csr);
}

/* ---- multithreaded driver for the CPU baseline (channels split evenly) ---- */
typedef struct { Biquad *bq; sample_t *s; uint32_t c0, c1, T, packet; } mt_job;
static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    ref_eq_many(j->bq + (size_t)j->c0 * MAX_BANDS, j->s + (size_t)j->c0 * j->T, j->c1 - j->c0, j->T, j->packet);
    return NULL;
}
/* returns elapsed seconds (CLOCK_MONOTONIC) for one pass over [C][T] */
double ref_eq_many_mt(Biquad *bq, sample_t *samples, uint32_t C, uint32_t T, uint32_t packet, uint32_t nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    mt_job jobs[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t i = 0; i < nthreads; i++) {
        jobs[i] = (mt_job){ bq, samples, (uint32_t)((uint64_t)C * i / nthreads), (uint32_t)((uint64_t)C * (i + 1) / nthreads), T, packet };
        pthread_create(&th[i], NULL, mt_worker, &jobs[i]);
    }
    for (uint32_t i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_xfeed_coeffs(CrossfeedState *st, uint8_t enabled, uint8_t itd, uint8_t preset, float fc, float feed_db, float fs)
{
    CrossfeedConfig cfg = { .enabled = enabled, .itd_enabled = itd, .preset = preset, .custom_fc = fc, .custom_feed_db = feed_db };
    crossfeed_compute_coefficients(st, &cfg, fs);
}
void ref_xfeed(CrossfeedState *st, sample_t *l, sample_t *r, uint32_t count)
{
    unsigned csr = ftz_enter();
    for (uint32_t i = 0; i < count; i++) crossfeed_process_stereo(st, &l[i], &r[i]);
    ftz_leave(csr);
}

void ref_lev_coeffs(LevellerCoeffs *out, float amount, uint8_t speed, float max_gain_db, float gate_db, float fs)
{
    LevellerConfig cfg = { .enabled = true, .amount = amount, .speed = speed, .max_gain_db = max_gain_db,
                           .lookahead = true, .gate_threshold_db = gate_db };
    leveller_compute_coefficients(out, &cfg, fs);
}
void ref_lev_reset(LevellerState *st) { leveller_reset_state(st); }
void ref_leveller(LevellerState *st, const LevellerCoeffs *c, int lookahead, sample_t *l, sample_t *r, uint32_t count)
{
    LevellerConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.enabled = true;
    cfg.lookahead = lookahead != 0;
    unsigned csr = ftz_enter();
    leveller_process_block(st, c, &cfg, l, r, count);
    ftz_leave(csr);
}

/* copies the freshly computed active table [61][2] to `out` */
void ref_loud_table(LoudnessCoeffs *out, float ref_spl, float intensity_pct, float fs)
{
    loudness_recompute_table(ref_spl, intensity_pct, fs);
    memcpy(out, loudness_active_table, sizeof(LoudnessCoeffs) * LOUDNESS_VOL_STEPS * LOUDNESS_BIQUAD_COUNT);
}

/* dsp_update_delay_samples() works on globals; expose it for one row */
int32_t ref_delay_samples(float delay_ms, float fs, int is_last)
{
    memset(channel_delays_ms, 0, sizeof(float) * NUM_CHANNELS);
    int out = is_last ? (NUM_DELAY_CHANNELS - 1) : 0;
    channel_delays_ms[CH_OUT_1 + out] = delay_ms;
    dsp_update_delay_samples(fs);
    return channel_delay_samples[out];
}
