/*
 * orc_params.c — parameter -> coefficient restatements (TEST INFRASTRUCTURE).
 * Float math through glibc libm, translation unit built with
 * -ffp-contract=off so results are bit-identical to the reference sources
 * compiled the same way (oracle/_ref/libdspi_ref_f32_strict.so, ..._q28.so).
 */
#include <math.h>
#include <string.h>
#include "dspi_oracle.h"
#include "orc_internal.h"

#define ORC_PI 3.1415926535f

/* dsp_pipeline.c:6-17 */
static int eq_is_flat(const orc_eq_param *p)
{
    if (p->type == ORC_FLAT) return 1;
    if (p->freq <= 0.0f) return 1;
    if (p->type == ORC_PEAKING || p->type == ORC_LOWSHELF || p->type == ORC_HIGHSHELF)
        if (fabsf(p->gain_db) < 0.01f) return 1;
    return 0;
}

/* RBJ cookbook section shared by both stores, dsp_pipeline.c:145-156 */
static void rbj(const orc_eq_param *p, float A, float fs, float *b, float *a)
{
    float omega = 2.0f * ORC_PI * p->freq / fs;
    float sn = orc_sinf(omega), cs = orc_cosf(omega);
    float alpha = sn / (2.0f * p->Q);
    float a0 = 1.0f, a1 = 0.0f, a2 = 0.0f, b0 = 1.0f, b1 = 0.0f, b2 = 0.0f;
    switch (p->type) {
    case ORC_LOWPASS:  b0 = (1 - cs) / 2; b1 = 1 - cs;    b2 = (1 - cs) / 2; a0 = 1 + alpha; a1 = -2 * cs; a2 = 1 - alpha; break;
    case ORC_HIGHPASS: b0 = (1 + cs) / 2; b1 = -(1 + cs); b2 = (1 + cs) / 2; a0 = 1 + alpha; a1 = -2 * cs; a2 = 1 - alpha; break;
    case ORC_PEAKING:  b0 = 1 + alpha * A; b1 = -2 * cs; b2 = 1 - alpha * A; a0 = 1 + alpha / A; a1 = -2 * cs; a2 = 1 - alpha / A; break;
    case ORC_LOWSHELF:
        b0 = A * ((A + 1) - (A - 1) * cs + 2 * sqrtf(A) * alpha);
        b1 = 2 * A * ((A - 1) - (A + 1) * cs);
        b2 = A * ((A + 1) - (A - 1) * cs - 2 * sqrtf(A) * alpha);
        a0 = (A + 1) + (A - 1) * cs + 2 * sqrtf(A) * alpha;
        a1 = -2 * ((A - 1) + (A + 1) * cs);
        a2 = (A + 1) + (A - 1) * cs - 2 * sqrtf(A) * alpha;
        break;
    case ORC_HIGHSHELF:
        b0 = A * ((A + 1) + (A - 1) * cs + 2 * sqrtf(A) * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * cs);
        b2 = A * ((A + 1) + (A - 1) * cs - 2 * sqrtf(A) * alpha);
        a0 = (A + 1) - (A - 1) * cs + 2 * sqrtf(A) * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * cs);
        a2 = (A + 1) - (A - 1) * cs - 2 * sqrtf(A) * alpha;
        break;
    default: break;
    }
    b[0] = b0; b[1] = b1; b[2] = b2; a[0] = a0; a[1] = a1; a[2] = a2;
}

/* dsp_pipeline.c:78-81 — clamps are written back into the recipe (quirk 8) */
static void eq_clamp(orc_eq_param *p, float fs)
{
    if (p->Q < 0.1f) p->Q = 0.1f;
    if (p->Q > 20.0f) p->Q = 20.0f;
    if (p->freq < 10.0f) p->freq = 10.0f;
    if (p->freq > fs * 0.45f) p->freq = fs * 0.45f;
}

/* dsp_pipeline.c:61-175, PICO_RP2350 branch */
void orc_eq_coeffs_f32(orc_eq_param *p, orc_biquad_f32 *bq, float fs)
{
    if (eq_is_flat(p) || fs == 0) {                                      /* :62-73 */
        bq->bypass = 1;
        bq->b0 = 1.0f; bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0.0f;
        bq->sva1 = bq->sva2 = bq->sva3 = 0.0f;
        bq->svm0 = bq->svm1 = bq->svm2 = 0.0f;
        bq->use_svf = 0;
        return;
    }
    bq->bypass = 0;
    eq_clamp(p, fs);
    float A = orc_powf(10.0f, p->gain_db / 40.0f);                           /* :83 */
    uint8_t was_svf = bq->use_svf;                                       /* :87-92 */
    bq->use_svf = (p->freq < (fs / 7.5f));
    if (was_svf != bq->use_svf) { bq->s1 = bq->s2 = 0.0f; bq->svic1eq = bq->svic2eq = 0.0f; }

    if (bq->use_svf) {                                                   /* :94-138 */
        float g = orc_tanf(ORC_PI * p->freq / fs);
        float k = 1.0f / p->Q;
        switch (p->type) {
        case ORC_PEAKING:   k = 1.0f / (p->Q * A); break;
        case ORC_LOWSHELF:  g = g / sqrtf(A); break;
        case ORC_HIGHSHELF: g = g * sqrtf(A); break;
        default: break;
        }
        float a1 = 1.0f / (1.0f + g * (g + k));
        float a2 = g * a1;
        float a3 = g * a2;
        float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f;
        switch (p->type) {
        case ORC_LOWPASS:   m0 = 0.0f;  m1 = 0.0f;                 m2 = 1.0f;          break;
        case ORC_HIGHPASS:  m0 = 1.0f;  m1 = -k;                   m2 = -1.0f;         break;
        case ORC_PEAKING:   m0 = 1.0f;  m1 = k * (A * A - 1.0f);   m2 = 0.0f;          break;
        case ORC_LOWSHELF:  m0 = 1.0f;  m1 = k * (A - 1.0f);       m2 = A * A - 1.0f;  break;
        case ORC_HIGHSHELF: m0 = A * A; m1 = k * (1.0f - A) * A;   m2 = 1.0f - A * A;  break;
        default: break;
        }
        bq->sva1 = a1; bq->sva2 = a2; bq->sva3 = a3;
        bq->svm0 = m0; bq->svm1 = m1; bq->svm2 = m2;
        bq->svf_type = p->type;
        bq->b0 = 1.0f; bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0.0f;        /* :136 */
        return;
    }
    bq->sva1 = bq->sva2 = bq->sva3 = 0.0f;                               /* :141-142 */
    bq->svm0 = bq->svm1 = bq->svm2 = 0.0f;
    float b[3], a[3];
    rbj(p, A, fs, b, a);
    float inv_a0 = 1.0f / a[0];                                          /* :160-165 */
    bq->b0 = b[0] * inv_a0; bq->b1 = b[1] * inv_a0; bq->b2 = b[2] * inv_a0;
    bq->a1 = a[1] * inv_a0; bq->a2 = a[2] * inv_a0;
}

/* dsp_pipeline.c:61-175, RP2040 branch: truncating Q28 store :168-173 */
void orc_eq_coeffs_q28(orc_eq_param *p, orc_biquad_q28 *bq, float fs)
{
    if (eq_is_flat(p) || fs == 0) {
        bq->bypass = 1;
        bq->b0 = 1 << 28; bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0;
        return;
    }
    bq->bypass = 0;
    eq_clamp(p, fs);
    float A = orc_powf(10.0f, p->gain_db / 40.0f);
    float b[3], a[3];
    rbj(p, A, fs, b, a);
    float scale = (float)(1LL << 28);
    bq->b0 = orc_f2i_sat((b[0] / a[0]) * scale);
    bq->b1 = orc_f2i_sat((b[1] / a[0]) * scale);
    bq->b2 = orc_f2i_sat((b[2] / a[0]) * scale);
    bq->a1 = orc_f2i_sat((a[1] / a[0]) * scale);
    bq->a2 = orc_f2i_sat((a[2] / a[0]) * scale);
}

/* crossfeed.c:25-29, 35-127 */
static int xfeed_core(const orc_xfeed_cfg *cfg, float fs, float *lp_a0, float *lp_b1, float *ap_a)
{
    static const float presets[3][2] = { {700.0f, 4.5f}, {700.0f, 6.0f}, {650.0f, 9.5f} };
    if (!cfg->enabled || fs < 1.0f) return 0;
    float fc, feed_db;
    if (cfg->preset < 3) { fc = presets[cfg->preset][0]; feed_db = presets[cfg->preset][1]; }
    else {
        fc = cfg->custom_fc; feed_db = cfg->custom_feed_db;
        if (fc < 500.0f) fc = 500.0f;
        if (fc > 2000.0f) fc = 2000.0f;
        if (feed_db < 0.0f) feed_db = 0.0f;
        if (feed_db > 15.0f) feed_db = 15.0f;
    }
    float level_ratio = orc_powf(10.0f, feed_db / 20.0f);                    /* :67 */
    float G = 1.0f / (1.0f + level_ratio);
    float x = orc_expf(-2.0f * ORC_PI * fc / fs);                            /* :75 */
    *lp_a0 = G * (1.0f - x);
    *lp_b1 = x;
    if (cfg->itd_enabled) {                                              /* :98-109 */
        float lp_delay_sec = x / ((1.0f - x) * fs);
        float remaining = 0.000220f - lp_delay_sec;
        if (remaining > 0.0f) { float D = remaining * fs; *ap_a = (1.0f - D) / (1.0f + D); }
        else *ap_a = 1.0f;
    } else *ap_a = 1.0f;
    return 1;
}

void orc_xfeed_coeffs_f32(orc_xfeed_f32 *st, const orc_xfeed_cfg *cfg, float fs)
{
    float a0, b1, ap;
    memset(st, 0, sizeof(*st));
    if (!xfeed_core(cfg, fs, &a0, &b1, &ap)) return;
    st->lp_a0 = a0; st->lp_b1 = b1; st->ap_a = ap;
}

void orc_xfeed_coeffs_q28(orc_xfeed_q28 *st, const orc_xfeed_cfg *cfg, float fs)
{
    float a0, b1, ap;
    memset(st, 0, sizeof(*st));
    if (!xfeed_core(cfg, fs, &a0, &b1, &ap)) return;
    float scale = (float)(1LL << 28);                                    /* :116-119 */
    st->lp_a0 = orc_f2i_sat(a0 * scale);
    st->lp_b1 = orc_f2i_sat(b1 * scale);
    st->ap_a  = orc_f2i_sat(ap * scale);
}

/* leveller.c:23-27, 37-40, 42-89 */
static float lev_alpha(float fs, float t)
{
    if (t <= 0.0f || fs <= 0.0f) return 0.0f;
    return orc_expf(-orc_logf(10.0f) / (fs * t));
}

void orc_lev_coeffs_compute(orc_lev_coeffs *out, const orc_lev_cfg *cfg, float fs)
{
    static const float speed[3][3] = { {0.100f, 2.000f, 0.400f}, {0.050f, 1.000f, 0.200f}, {0.020f, 0.500f, 0.100f} };
    if (fs < 1.0f) fs = 48000.0f;
    uint8_t spd = cfg->speed;
    if (spd >= 3) spd = 1;
    out->alpha_rms     = lev_alpha(fs, speed[spd][2]);
    out->alpha_attack  = lev_alpha(fs, speed[spd][0]);
    out->alpha_release = lev_alpha(fs, speed[spd][1]);
    out->threshold_db  = -20.0f;
    out->knee_width_db = 6.0f;
    float gate = cfg->gate_threshold_db;
    if (gate < -96.0f) gate = -96.0f;
    if (gate > 0.0f) gate = 0.0f;
    out->gate_threshold_db = gate;
    float amount = cfg->amount;
    if (amount < 0.0f) amount = 0.0f;
    if (amount > 100.0f) amount = 100.0f;
    float norm = amount / 100.0f;
    out->ratio = 1.0f + norm * 19.0f;
    float max_g = cfg->max_gain_db;
    if (max_g < 0.0f) max_g = 0.0f;
    if (max_g > 35.0f) max_g = 35.0f;
    out->max_gain_db = max_g;
    out->makeup_db = 0.0f;
}

void orc_lev_reset_f32(orc_lev_state_f32 *st)
{
    memset(st, 0, sizeof(*st));
    st->gain_linear = 1.0f;
    st->gain_prev_linear = 1.0f;
}
void orc_lev_reset_q28(orc_lev_state_q28 *st)
{
    memset(st, 0, sizeof(*st));
    st->gain_q28 = 1 << 28;
    st->gain_prev_q28 = 1 << 28;
}

/* loudness.c:37-50 */
static float iso226_spl(float Tf, float af, float Lu, float phon)
{
    float B = 0.4f * orc_powf(10.0f, (Tf + Lu) / 10.0f - 9.0f);
    float threshold = orc_powf(B, af);
    float Af = 4.47e-3f * (orc_powf(10.0f, 0.025f * phon) - 1.15f) + threshold;
    if (Af < 1e-10f) Af = 1e-10f;
    return (10.0f / af) * orc_log10f(Af) - Lu + 94.0f;
}
/* loudness.c:54-78 */
static float loud_comp_db(float Tf, float af, float Lu, float ref_spl, float eff_phon, float intensity)
{
    if (eff_phon >= ref_spl) return 0.0f;
    float spl_ref = iso226_spl(Tf, af, Lu, ref_spl);
    float spl_eff = iso226_spl(Tf, af, Lu, eff_phon);
    float flat_change = eff_phon - ref_spl;
    float freq_change = spl_eff - spl_ref;
    float comp = freq_change - flat_change;
    comp *= (intensity / 100.0f);
    return comp;
}
/* loudness.c:85-163, float branch */
static void shelf_f32(float freq, float Q, float gain_db, int high, float fs, orc_loud_f32 *o)
{
    if (fabsf(gain_db) < 0.01f) {
        o->bypass = 1;
        o->sva1 = o->sva2 = o->sva3 = 0.0f;
        o->svm0 = o->svm1 = o->svm2 = 0.0f;
        return;
    }
    o->bypass = 0;
    float A = orc_powf(10.0f, gain_db / 40.0f);
    float g = orc_tanf(ORC_PI * freq / fs);
    float sqrtA = sqrtf(A);
    if (high) g = g * sqrtA; else g = g / sqrtA;
    float k = 1.0f / Q;
    o->sva1 = 1.0f / (1.0f + g * (g + k));
    o->sva2 = g * o->sva1;
    o->sva3 = g * o->sva2;
    if (high) { o->svm0 = A * A; o->svm1 = k * (1.0f - A) * A; o->svm2 = 1.0f - A * A; }
    else      { o->svm0 = 1.0f;  o->svm1 = k * (A - 1.0f);     o->svm2 = A * A - 1.0f; }
}
/* loudness.c:85-163, Q28 branch */
static void shelf_q28(float freq, float Q, float gain_db, int high, float fs, orc_loud_q28 *o)
{
    if (fabsf(gain_db) < 0.01f) {
        o->bypass = 1;
        o->b0 = 1 << 28; o->b1 = o->b2 = o->a1 = o->a2 = 0;
        return;
    }
    o->bypass = 0;
    float A = orc_powf(10.0f, gain_db / 40.0f);
    float omega = 2.0f * ORC_PI * freq / fs;
    float sn = orc_sinf(omega), cs = orc_cosf(omega);
    float alpha = sn / (2.0f * Q);
    float sqrtA = sqrtf(A);
    float a0, a1, a2, b0, b1, b2;
    if (high) {
        b0 = A * ((A + 1) + (A - 1) * cs + 2 * sqrtA * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * cs);
        b2 = A * ((A + 1) + (A - 1) * cs - 2 * sqrtA * alpha);
        a0 = (A + 1) - (A - 1) * cs + 2 * sqrtA * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * cs);
        a2 = (A + 1) - (A - 1) * cs - 2 * sqrtA * alpha;
    } else {
        b0 = A * ((A + 1) - (A - 1) * cs + 2 * sqrtA * alpha);
        b1 = 2 * A * ((A - 1) - (A + 1) * cs);
        b2 = A * ((A + 1) - (A - 1) * cs - 2 * sqrtA * alpha);
        a0 = (A + 1) + (A - 1) * cs + 2 * sqrtA * alpha;
        a1 = -2 * ((A - 1) + (A + 1) * cs);
        a2 = (A + 1) + (A - 1) * cs - 2 * sqrtA * alpha;
    }
    float scale = (float)(1LL << 28);
    o->b0 = orc_f2i_sat((b0 / a0) * scale);
    o->b1 = orc_f2i_sat((b1 / a0) * scale);
    o->b2 = orc_f2i_sat((b2 / a0) * scale);
    o->a1 = orc_f2i_sat((a1 / a0) * scale);
    o->a2 = orc_f2i_sat((a2 / a0) * scale);
}

/* loudness.c:169-217 (one table; the double buffering is control plane) */
static void loud_gains(int idx, float ref_spl, float intensity, float *low_db, float *high_db)
{
    float vol_db = (float)(idx - 60);
    float eff = ref_spl + vol_db;
    if (eff < 20.0f) eff = 20.0f;
    if (eff > ref_spl) eff = ref_spl;
    *low_db  = loud_comp_db(44.0f, 0.432f, 80.4f, ref_spl, eff, intensity);   /* ISO 226 @50 Hz, :20-22 */
    *high_db = loud_comp_db(13.9f, 0.301f, 17.8f, ref_spl, eff, intensity);   /* ISO 226 @10 kHz, :26-28 */
}

void orc_loud_table_f32(orc_loud_f32 table[ORC_LOUD_STEPS][2], float ref_spl, float intensity, float fs)
{
    if (fs < 1.0f) fs = 48000.0f;
    if (ref_spl < 40.0f) ref_spl = 40.0f;
    if (ref_spl > 100.0f) ref_spl = 100.0f;
    memset(table, 0, sizeof(orc_loud_f32) * ORC_LOUD_STEPS * 2);
    for (int i = 0; i < ORC_LOUD_STEPS; i++) {
        float lo, hi;
        loud_gains(i, ref_spl, intensity, &lo, &hi);
        shelf_f32(200.0f, 0.707f, lo, 0, fs, &table[i][0]);
        shelf_f32(6000.0f, 0.707f, hi, 1, fs, &table[i][1]);
    }
}

void orc_loud_table_q28(orc_loud_q28 table[ORC_LOUD_STEPS][2], float ref_spl, float intensity, float fs)
{
    if (fs < 1.0f) fs = 48000.0f;
    if (ref_spl < 40.0f) ref_spl = 40.0f;
    if (ref_spl > 100.0f) ref_spl = 100.0f;
    memset(table, 0, sizeof(orc_loud_q28) * ORC_LOUD_STEPS * 2);
    for (int i = 0; i < ORC_LOUD_STEPS; i++) {
        float lo, hi;
        loud_gains(i, ref_spl, intensity, &lo, &hi);
        shelf_q28(200.0f, 0.707f, lo, 0, fs, &table[i][0]);
        shelf_q28(6000.0f, 0.707f, hi, 1, fs, &table[i][1]);
    }
}

/* dsp_pipeline.c:216-239 — one output; `is_last` adds SUB_ALIGN_SAMPLES (128) */
int32_t orc_delay_samples(float delay_ms, float fs, int is_last, int32_t max_delay)
{
    if (is_last) {
        float align_ms = (float)128 / fs * 1000.0f;
        delay_ms += align_ms;
    }
    int32_t s = orc_f2i_sat(delay_ms * fs / 1000.0f);
    if (s > max_delay) s = max_delay;
    if (s < 0) s = 0;
    return s;
}

/* usb_audio.c:410-440 — db_to_vol table lookup; result stored in an int16 */
int16_t orc_host_vol_mul(int16_t volume, uint8_t *vol_index_out)
{
    static const uint16_t db_to_vol[61] = {
        0x0000, 0x0025, 0x0029, 0x002e, 0x0034, 0x003a, 0x0041, 0x0049,
        0x0052, 0x005c, 0x0068, 0x0074, 0x0082, 0x0092, 0x00a4, 0x00b8,
        0x00cf, 0x00e8, 0x0104, 0x0124, 0x0148, 0x0170, 0x019d, 0x01cf,
        0x0207, 0x0247, 0x028e, 0x02de, 0x0337, 0x039c, 0x040c, 0x048b,
        0x0519, 0x05b8, 0x066a, 0x0733, 0x0814, 0x0910, 0x0a2b, 0x0b68,
        0x0ccd, 0x0e5d, 0x101d, 0x1215, 0x1449, 0x16c3, 0x198a, 0x1ca8,
        0x2027, 0x2413, 0x287a, 0x2d6b, 0x32f5, 0x392d, 0x4027, 0x47fb,
        0x50c3, 0x5a9e, 0x65ad, 0x7215, 0x8000
    };
    int16_t v = (int16_t)(volume + 60 * 256);
    if (v < 0) v = 0;
    if (v >= 61 * 256) v = 61 * 256 - 1;
    uint8_t idx = (uint8_t)(((uint16_t)v) >> 8);
    if (vol_index_out) *vol_index_out = idx;
    return (int16_t)db_to_vol[idx];
}

/* usb_audio.c:456-498 — preset-mute envelope, advanced once per packet */
float orc_mute_envelope(uint8_t *preset_loading, uint32_t *preset_mute_counter, float *smooth_gain,
                        uint32_t sample_count, uint32_t sample_rate_hz)
{
    const int mute_active_for_packet = *preset_loading != 0;                      /* :469 */
    if (mute_active_for_packet) {
        if (*preset_mute_counter > sample_count) *preset_mute_counter -= sample_count;   /* :472-473 */
        else { *preset_mute_counter = 0; *preset_loading = 0; }                   /* :475-476 */
    }
    const float target = mute_active_for_packet ? 0.0f : 1.0f;                    /* :480 */
    if (sample_count == 0) { *smooth_gain = target; return target; }              /* :481-484 */
    uint64_t ts = ((uint64_t)sample_rate_hz * 8u + 999u) / 1000u;                 /* :459-464, PRESET_MUTE_TRANSITION_MS = 8 */
    if (ts < 1u) ts = 1u;
    if (ts > 0xFFFFFFFFu) ts = 0xFFFFFFFFu;
    float step = (float)sample_count / (float)(uint32_t)ts;                       /* :486 */
    if (step > 1.0f) step = 1.0f;
    float g = *smooth_gain;
    if (g < target)      { g += step; if (g > target) g = target; }               /* :489-491 */
    else if (g > target) { g -= step; if (g < target) g = target; }               /* :492-494 */
    *smooth_gain = g;
    return g;
}
