/*
 * host_params.c — host side of the parameter API: EqParamPacket -> Biquad.
 *
 * Takes the role of dsp_compute_coefficients() (reference
 * firmware/DSPi/dsp_pipeline.c:61-175) for both coefficient stores.  Runs on
 * the host so that coefficients are produced with the host libm and handed to
 * the GPU as data.  Built with -ffp-contract=off: every
 * operation is rounded on its own, the same arithmetic as the reference
 * compiled without contraction.
 */
#include <math.h>
#include <string.h>
#include "dspi_b200.h"

#define PI_F 3.1415926535f          /* the reference's literal, dsp_pipeline.c:97,145 */

/* dsp_pipeline.c:6-17 */
static int recipe_is_flat(const dspi_eq_param *p)
{
    if (p->type == DSPI_FILTER_FLAT || p->freq <= 0.0f) return 1;
    switch (p->type) {
    case DSPI_FILTER_PEAKING: case DSPI_FILTER_LOWSHELF: case DSPI_FILTER_HIGHSHELF:
        return fabsf(p->gain_db) < 0.01f;
    default:
        return 0;
    }
}

/* dsp_pipeline.c:78-81 — the clamps are written back into the caller's recipe */
static void recipe_clamp(dspi_eq_param *p, float fs)
{
    float q = p->Q, f = p->freq;
    if (q < 0.1f) q = 0.1f;
    if (q > 20.0f) q = 20.0f;
    if (f < 10.0f) f = 10.0f;
    if (f > fs * 0.45f) f = fs * 0.45f;
    p->Q = q;
    p->freq = f;
}

/* un-normalised RBJ cookbook section, dsp_pipeline.c:145-156; n = {b0,b1,b2}, d = {a0,a1,a2} */
static void cookbook(const dspi_eq_param *p, float A, float fs, float n[3], float d[3])
{
    const float omega = 2.0f * PI_F * p->freq / fs;
    const float sn = sinf(omega), cs = cosf(omega);
    const float alpha = sn / (2.0f * p->Q);
    n[0] = 1.0f; n[1] = 0.0f; n[2] = 0.0f;
    d[0] = 1.0f; d[1] = 0.0f; d[2] = 0.0f;
    switch (p->type) {
    case DSPI_FILTER_LOWPASS:
        n[0] = (1 - cs) / 2; n[1] = 1 - cs; n[2] = (1 - cs) / 2;
        d[0] = 1 + alpha; d[1] = -2 * cs; d[2] = 1 - alpha;
        break;
    case DSPI_FILTER_HIGHPASS:
        n[0] = (1 + cs) / 2; n[1] = -(1 + cs); n[2] = (1 + cs) / 2;
        d[0] = 1 + alpha; d[1] = -2 * cs; d[2] = 1 - alpha;
        break;
    case DSPI_FILTER_PEAKING:
        n[0] = 1 + alpha * A; n[1] = -2 * cs; n[2] = 1 - alpha * A;
        d[0] = 1 + alpha / A; d[1] = -2 * cs; d[2] = 1 - alpha / A;
        break;
    case DSPI_FILTER_LOWSHELF:
        n[0] = A * ((A + 1) - (A - 1) * cs + 2 * sqrtf(A) * alpha);
        n[1] = 2 * A * ((A - 1) - (A + 1) * cs);
        n[2] = A * ((A + 1) - (A - 1) * cs - 2 * sqrtf(A) * alpha);
        d[0] = (A + 1) + (A - 1) * cs + 2 * sqrtf(A) * alpha;
        d[1] = -2 * ((A - 1) + (A + 1) * cs);
        d[2] = (A + 1) + (A - 1) * cs - 2 * sqrtf(A) * alpha;
        break;
    case DSPI_FILTER_HIGHSHELF:
        n[0] = A * ((A + 1) + (A - 1) * cs + 2 * sqrtf(A) * alpha);
        n[1] = -2 * A * ((A - 1) + (A + 1) * cs);
        n[2] = A * ((A + 1) + (A - 1) * cs - 2 * sqrtf(A) * alpha);
        d[0] = (A + 1) - (A - 1) * cs + 2 * sqrtf(A) * alpha;
        d[1] = 2 * ((A - 1) - (A + 1) * cs);
        d[2] = (A + 1) - (A - 1) * cs - 2 * sqrtf(A) * alpha;
        break;
    default:
        break;
    }
}

void dspi_compute_coefficients_f32(dspi_eq_param *p, dspi_biquad_f32 *bq, float fs)
{
    if (recipe_is_flat(p) || fs == 0) {                 /* dsp_pipeline.c:62-73 */
        bq->bypass = 1;
        bq->use_svf = 0;
        bq->b0 = 1.0f;
        bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0.0f;
        bq->sva1 = bq->sva2 = bq->sva3 = 0.0f;
        bq->svm0 = bq->svm1 = bq->svm2 = 0.0f;
        return;
    }
    bq->bypass = 0;
    recipe_clamp(p, fs);
    const float A = powf(10.0f, p->gain_db / 40.0f);    /* :83 */

    /* :87-92 — pick the topology; a flip clears both state pairs */
    const uint8_t svf_now = (p->freq < (fs / 7.5f)) ? 1 : 0;
    if (svf_now != bq->use_svf) {
        bq->s1 = bq->s2 = 0.0f;
        bq->svic1eq = bq->svic2eq = 0.0f;
    }
    bq->use_svf = svf_now;

    if (svf_now) {                                      /* :94-138 Cytomic SVF */
        float g = tanf(PI_F * p->freq / fs);
        float k = 1.0f / p->Q;
        if (p->type == DSPI_FILTER_PEAKING)        k = 1.0f / (p->Q * A);
        else if (p->type == DSPI_FILTER_LOWSHELF)  g = g / sqrtf(A);
        else if (p->type == DSPI_FILTER_HIGHSHELF) g = g * sqrtf(A);
        const float c1 = 1.0f / (1.0f + g * (g + k));
        const float c2 = g * c1;
        const float c3 = g * c2;
        float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f;
        switch (p->type) {
        case DSPI_FILTER_LOWPASS:   m2 = 1.0f; break;
        case DSPI_FILTER_HIGHPASS:  m0 = 1.0f; m1 = -k; m2 = -1.0f; break;
        case DSPI_FILTER_PEAKING:   m0 = 1.0f; m1 = k * (A * A - 1.0f); break;
        case DSPI_FILTER_LOWSHELF:  m0 = 1.0f; m1 = k * (A - 1.0f); m2 = A * A - 1.0f; break;
        case DSPI_FILTER_HIGHSHELF: m0 = A * A; m1 = k * (1.0f - A) * A; m2 = 1.0f - A * A; break;
        default: break;
        }
        bq->sva1 = c1; bq->sva2 = c2; bq->sva3 = c3;
        bq->svm0 = m0; bq->svm1 = m1; bq->svm2 = m2;
        bq->svf_type = p->type;
        bq->b0 = 1.0f;
        bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0.0f;
        return;
    }

    bq->sva1 = bq->sva2 = bq->sva3 = 0.0f;              /* :141-142 */
    bq->svm0 = bq->svm1 = bq->svm2 = 0.0f;
    float n[3], d[3];
    cookbook(p, A, fs, n, d);
    const float inv_a0 = 1.0f / d[0];                   /* :160-165 */
    bq->b0 = n[0] * inv_a0;
    bq->b1 = n[1] * inv_a0;
    bq->b2 = n[2] * inv_a0;
    bq->a1 = d[1] * inv_a0;
    bq->a2 = d[2] * inv_a0;
}

/* (int32_t) cast with the firmware's saturating semantics (soft-float
 * __aeabi_f2iz on the RP2040); in range for every valid coefficient. */
static int32_t to_q28(float v)
{
    const float x = v * (float)(1LL << 28);
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

void dspi_compute_coefficients_q28(dspi_eq_param *p, dspi_biquad_q28 *bq, float fs)
{
    if (recipe_is_flat(p) || fs == 0) {
        bq->bypass = 1;
        bq->b0 = 1 << 28;
        bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0;
        return;
    }
    bq->bypass = 0;
    recipe_clamp(p, fs);
    const float A = powf(10.0f, p->gain_db / 40.0f);
    float n[3], d[3];
    cookbook(p, A, fs, n, d);
    bq->b0 = to_q28(n[0] / d[0]);                       /* :168-173 truncating store */
    bq->b1 = to_q28(n[1] / d[0]);
    bq->b2 = to_q28(n[2] / d[0]);
    bq->a1 = to_q28(d[1] / d[0]);
    bq->a2 = to_q28(d[2] / d[0]);
}

/* ------------------------------------------------------------------------------------------------
 * Crossfeed, leveller, loudness and host-volume parameter functions (float / RP2350 stores).
 * Same role as crossfeed_compute_coefficients() (crossfeed.c:35-127),
 * leveller_compute_coefficients() (leveller.c:42-89), loudness_recompute_table()
 * (loudness.c:169-217) and audio_set_volume() (usb_audio.c:410-440).
 * ---------------------------------------------------------------------------------------------- */

void dspi_crossfeed_compute_coefficients_f32(dspi_crossfeed_state_f32 *st, const dspi_crossfeed_config *cfg, float fs)
{
    static const float preset_fc[3] = { 700.0f, 700.0f, 650.0f };          /* crossfeed.c:25-29 */
    static const float preset_db[3] = { 4.5f, 6.0f, 9.5f };
    memset(st, 0, sizeof(*st));                                             /* disabled: crossfeed_init() */
    if (!cfg->enabled || fs < 1.0f) return;
    float fc, feed_db;
    if (cfg->preset < 3) {
        fc = preset_fc[cfg->preset];
        feed_db = preset_db[cfg->preset];
    } else {                                                                /* :47-52 custom, clamped */
        fc = cfg->custom_fc;
        feed_db = cfg->custom_feed_db;
        if (fc < 500.0f) fc = 500.0f;
        if (fc > 2000.0f) fc = 2000.0f;
        if (feed_db < 0.0f) feed_db = 0.0f;
        if (feed_db > 15.0f) feed_db = 15.0f;
    }
    const float level_ratio = powf(10.0f, feed_db / 20.0f);                 /* :67-68 */
    const float G = 1.0f / (1.0f + level_ratio);
    const float x = expf(-2.0f * PI_F * fc / fs);                           /* :75-77 */
    st->lp_a0 = G * (1.0f - x);
    st->lp_b1 = x;
    float ap = 1.0f;                                                        /* :97-109 */
    if (cfg->itd_enabled) {
        const float lp_delay_sec = x / ((1.0f - x) * fs);
        const float remaining_sec = 0.000220f - lp_delay_sec;
        if (remaining_sec > 0.0f) {
            const float D = remaining_sec * fs;
            ap = (1.0f - D) / (1.0f + D);
        }
    }
    st->ap_a = ap;
}

static float retention(float fs, float seconds)                             /* leveller.c:37-40 */
{
    if (seconds <= 0.0f || fs <= 0.0f) return 0.0f;
    return expf(-logf(10.0f) / (fs * seconds));
}

void dspi_leveller_compute_coefficients(dspi_leveller_coeffs *out, const dspi_leveller_config *cfg, float fs)
{
    /* {attack, release, rms window} seconds per speed preset, leveller.c:23-27 */
    static const float presets[3][3] = { { 0.100f, 2.000f, 0.400f }, { 0.050f, 1.000f, 0.200f }, { 0.020f, 0.500f, 0.100f } };
    if (fs < 1.0f) fs = 48000.0f;
    const unsigned spd = cfg->speed >= 3 ? 1u : cfg->speed;
    out->alpha_rms = retention(fs, presets[spd][2]);
    out->alpha_attack = retention(fs, presets[spd][0]);
    out->alpha_release = retention(fs, presets[spd][1]);
    out->threshold_db = -20.0f;                                             /* leveller.h:51-52 */
    out->knee_width_db = 6.0f;
    float gate = cfg->gate_threshold_db;
    if (gate < -96.0f) gate = -96.0f;
    if (gate > 0.0f) gate = 0.0f;
    out->gate_threshold_db = gate;
    float amount = cfg->amount;
    if (amount < 0.0f) amount = 0.0f;
    if (amount > 100.0f) amount = 100.0f;
    out->ratio = 1.0f + (amount / 100.0f) * 19.0f;                          /* :76-77 */
    float max_g = cfg->max_gain_db;
    if (max_g < 0.0f) max_g = 0.0f;
    if (max_g > 35.0f) max_g = 35.0f;
    out->max_gain_db = max_g;
    out->makeup_db = 0.0f;
}

/* ISO 226:2003 equal-loudness SPL at one tabulated frequency, loudness.c:37-50 */
static float iso226(float Tf, float af, float Lu, float phon)
{
    const float B = 0.4f * powf(10.0f, (Tf + Lu) / 10.0f - 9.0f);
    float Af = 4.47e-3f * (powf(10.0f, 0.025f * phon) - 1.15f) + powf(B, af);
    if (Af < 1e-10f) Af = 1e-10f;
    return (10.0f / af) * log10f(Af) - Lu + 94.0f;
}

static float shelf_gain_db(float Tf, float af, float Lu, float ref_spl, float phon, float intensity_pct)   /* loudness.c:54-78 */
{
    if (phon >= ref_spl) return 0.0f;
    const float at_ref = iso226(Tf, af, Lu, ref_spl);
    const float at_eff = iso226(Tf, af, Lu, phon);
    const float flat_change = phon - ref_spl;
    const float freq_change = at_eff - at_ref;
    float comp = freq_change - flat_change;
    comp *= (intensity_pct / 100.0f);
    return comp;
}

static void shelf_svf(float freq, float Q, float gain_db, int high, float fs, dspi_loudness_coeffs_f32 *o)  /* loudness.c:85-130 */
{
    if (fabsf(gain_db) < 0.01f) {
        memset(o, 0, sizeof(*o));
        o->bypass = 1;
        return;
    }
    o->bypass = 0;
    const float A = powf(10.0f, gain_db / 40.0f);
    float g = tanf(PI_F * freq / fs);
    const float rootA = sqrtf(A);
    g = high ? g * rootA : g / rootA;
    const float k = 1.0f / Q;
    o->sva1 = 1.0f / (1.0f + g * (g + k));
    o->sva2 = g * o->sva1;
    o->sva3 = g * o->sva2;
    if (high) { o->svm0 = A * A; o->svm1 = k * (1.0f - A) * A; o->svm2 = 1.0f - A * A; }
    else { o->svm0 = 1.0f; o->svm1 = k * (A - 1.0f); o->svm2 = A * A - 1.0f; }
}

void dspi_loudness_compute_table_f32(dspi_loudness_coeffs_f32 table[61][2], float ref_spl, float intensity_pct, float fs)
{
    if (fs < 1.0f) fs = 48000.0f;
    if (ref_spl < 40.0f) ref_spl = 40.0f;
    if (ref_spl > 100.0f) ref_spl = 100.0f;
    memset(table, 0, sizeof(dspi_loudness_coeffs_f32) * 61 * 2);
    for (int step = 0; step < 61; step++) {
        float phon = ref_spl + (float)(step - 60);                          /* :186-191 */
        if (phon < 20.0f) phon = 20.0f;
        if (phon > ref_spl) phon = ref_spl;
        const float low_db = shelf_gain_db(44.0f, 0.432f, 80.4f, ref_spl, phon, intensity_pct);    /* 50 Hz row of ISO 226 Table 1 */
        const float high_db = shelf_gain_db(13.9f, 0.301f, 17.8f, ref_spl, phon, intensity_pct);   /* 10 kHz row */
        shelf_svf(200.0f, 0.707f, low_db, 0, fs, &table[step][0]);
        shelf_svf(6000.0f, 0.707f, high_db, 1, fs, &table[step][1]);
    }
}

int16_t dspi_host_volume(int16_t volume_8_8, uint8_t *table_index)
{
    /* the UAC1 volume table of the firmware, usb_audio.c:410-420 (61 Q15 steps, -60 .. 0 dB) */
    static const uint16_t q15[61] = {
        0x0000, 0x0025, 0x0029, 0x002e, 0x0034, 0x003a, 0x0041, 0x0049, 0x0052, 0x005c, 0x0068, 0x0074, 0x0082, 0x0092, 0x00a4, 0x00b8,
        0x00cf, 0x00e8, 0x0104, 0x0124, 0x0148, 0x0170, 0x019d, 0x01cf, 0x0207, 0x0247, 0x028e, 0x02de, 0x0337, 0x039c, 0x040c, 0x048b,
        0x0519, 0x05b8, 0x066a, 0x0733, 0x0814, 0x0910, 0x0a2b, 0x0b68, 0x0ccd, 0x0e5d, 0x101d, 0x1215, 0x1449, 0x16c3, 0x198a, 0x1ca8,
        0x2027, 0x2413, 0x287a, 0x2d6b, 0x32f5, 0x392d, 0x4027, 0x47fb, 0x50c3, 0x5a9e, 0x65ad, 0x7215, 0x8000 };
    int16_t v = (int16_t)(volume_8_8 + 60 * 256);                           /* :430-432 */
    if (v < 0) v = 0;
    if (v >= 61 * 256) v = 61 * 256 - 1;
    const uint8_t idx = (uint8_t)(((uint16_t)v) >> 8);
    if (table_index) *table_index = idx;
    return (int16_t)q15[idx];                                               /* stored in an int16_t: 0 dB -> -32768 */
}

/* update_preamp(), usb_audio.c:244-250: dB -> the float and the Q28 gain the packet loop reads.  NaN / Inf are
 * rejected as there (returns -1, outputs untouched).  The Q28 conversion saturates like the Cortex-M VCVT. */
int dspi_preamp(float db, float *linear_out, int32_t *q28_out)
{
    if (!isfinite(db)) return -1;
    const float linear = powf(10.0f, db / 20.0f);
    const float scaled = linear * (float)(1 << 28);
    if (linear_out) *linear_out = linear;
    if (q28_out) *q28_out = scaled >= 2147483648.0f ? INT32_MAX : (int32_t)scaled;
    return 0;
}

/* update_master_volume(), usb_audio.c:255-269: clamp to [-128, 0] dB, -128 is the mute sentinel */
int dspi_master_volume(float db, float *linear_out, int32_t *q15_out)
{
    if (!isfinite(db)) return -1;
    if (db < -128.0f) db = -128.0f;
    if (db > 0.0f) db = 0.0f;
    float linear = 0.0f;
    int32_t q15 = 0;
    if (db > -128.0f) {
        linear = powf(10.0f, db / 20.0f);
        q15 = (int32_t)(linear * 32768.0f);
    }
    if (linear_out) *linear_out = linear;
    if (q15_out) *q15_out = q15;
    return 0;
}

/* update_preset_mute_envelope(), usb_audio.c:456-498, for one packet of `sample_count` frames: the state is
 * {preset_loading, preset_mute_counter (flash_storage.c:255-256), preset_mute_smooth_gain}; returns the gain
 * process_audio_packet() multiplies into the host volume.  dspi_preset_mute_arm() is what every flash-backed
 * operation does first (flash_storage.c:272-276, 347-348, 775-776): hold for max(512, 10 ms) samples. */
void dspi_preset_mute_arm(dspi_preset_mute *m, uint32_t sample_rate_hz)
{
    uint64_t samples = ((uint64_t)sample_rate_hz * 10u + 999u) / 1000u;
    if (samples < 512u) samples = 512u;
    m->counter = (uint32_t)samples;
    m->loading = 1;
}

float dspi_preset_mute_step(dspi_preset_mute *m, uint32_t sample_count, uint32_t sample_rate_hz)
{
    const int active = m->loading != 0;                                     /* :469 latched for this packet */
    if (active) {
        if (m->counter > sample_count) m->counter -= sample_count;
        else { m->counter = 0; m->loading = 0; }
    }
    const float target = active ? 0.0f : 1.0f;
    if (sample_count == 0) { m->smooth_gain = target; return target; }
    uint64_t ts = ((uint64_t)sample_rate_hz * 8u + 999u) / 1000u;           /* PRESET_MUTE_TRANSITION_MS 8 */
    if (ts < 1u) ts = 1u;
    if (ts > UINT32_MAX) ts = UINT32_MAX;
    float step = (float)sample_count / (float)(uint32_t)ts;
    if (step > 1.0f) step = 1.0f;
    float g = m->smooth_gain;
    if (g < target)      { g += step; if (g > target) g = target; }
    else if (g > target) { g -= step; if (g < target) g = target; }
    m->smooth_gain = g;
    return g;
}

/* ---- Q28 stores (RP2040 build of the same parameter functions) -------------------------------- */

void dspi_crossfeed_compute_coefficients_q28(dspi_crossfeed_state_q28 *st, const dspi_crossfeed_config *cfg, float fs)
{
    dspi_crossfeed_state_f32 f;
    dspi_crossfeed_compute_coefficients_f32(&f, cfg, fs);                   /* same float maths, crossfeed.c:35-109 */
    memset(st, 0, sizeof(*st));
    if (!cfg->enabled || fs < 1.0f) return;
    st->lp_a0 = to_q28(f.lp_a0);                                            /* :116-119 (int32_t)(x * 2^28) */
    st->lp_b1 = to_q28(f.lp_b1);
    st->ap_a = to_q28(f.ap_a);
}

static void shelf_rbj_q28(float freq, float Q, float gain_db, int high, float fs, dspi_loudness_coeffs_q28 *o)   /* loudness.c:85-162 */
{
    if (fabsf(gain_db) < 0.01f) {
        memset(o, 0, sizeof(*o));
        o->bypass = 1;
        o->b0 = 1 << 28;
        return;
    }
    o->bypass = 0;
    const float A = powf(10.0f, gain_db / 40.0f);
    const float omega = 2.0f * PI_F * freq / fs;
    const float sn = sinf(omega), cs = cosf(omega);
    const float alpha = sn / (2.0f * Q);
    const float rootA = sqrtf(A);
    float n0, n1, n2, d0, d1, d2;
    if (high) {
        n0 = A * ((A + 1) + (A - 1) * cs + 2 * rootA * alpha);
        n1 = -2 * A * ((A - 1) + (A + 1) * cs);
        n2 = A * ((A + 1) + (A - 1) * cs - 2 * rootA * alpha);
        d0 = (A + 1) - (A - 1) * cs + 2 * rootA * alpha;
        d1 = 2 * ((A - 1) - (A + 1) * cs);
        d2 = (A + 1) - (A - 1) * cs - 2 * rootA * alpha;
    } else {
        n0 = A * ((A + 1) - (A - 1) * cs + 2 * rootA * alpha);
        n1 = 2 * A * ((A - 1) - (A + 1) * cs);
        n2 = A * ((A + 1) - (A - 1) * cs - 2 * rootA * alpha);
        d0 = (A + 1) + (A - 1) * cs + 2 * rootA * alpha;
        d1 = -2 * ((A - 1) + (A + 1) * cs);
        d2 = (A + 1) + (A - 1) * cs - 2 * rootA * alpha;
    }
    o->b0 = to_q28(n0 / d0);
    o->b1 = to_q28(n1 / d0);
    o->b2 = to_q28(n2 / d0);
    o->a1 = to_q28(d1 / d0);
    o->a2 = to_q28(d2 / d0);
}

void dspi_loudness_compute_table_q28(dspi_loudness_coeffs_q28 table[61][2], float ref_spl, float intensity_pct, float fs)
{
    if (fs < 1.0f) fs = 48000.0f;
    if (ref_spl < 40.0f) ref_spl = 40.0f;
    if (ref_spl > 100.0f) ref_spl = 100.0f;
    memset(table, 0, sizeof(dspi_loudness_coeffs_q28) * 61 * 2);
    for (int step = 0; step < 61; step++) {
        float phon = ref_spl + (float)(step - 60);
        if (phon < 20.0f) phon = 20.0f;
        if (phon > ref_spl) phon = ref_spl;
        const float low_db = shelf_gain_db(44.0f, 0.432f, 80.4f, ref_spl, phon, intensity_pct);
        const float high_db = shelf_gain_db(13.9f, 0.301f, 17.8f, ref_spl, phon, intensity_pct);
        shelf_rbj_q28(200.0f, 0.707f, low_db, 0, fs, &table[step][0]);
        shelf_rbj_q28(6000.0f, 0.707f, high_db, 1, fs, &table[step][1]);
    }
}
