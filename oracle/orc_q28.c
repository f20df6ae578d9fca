/*
 * orc_q28.c — Q28 fixed-point oracle (RP2040 arithmetic) and the ΔΣ PDM core.
 * TEST INFRASTRUCTURE; see dspi_oracle.h.  Build with -fwrapv: the firmware
 * relies on 32-bit two's-complement wrap-around.
 */
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "dspi_oracle.h"
#include "orc_internal.h"

int orc_x86_cvt_mode = 0;
int orc_libm_f64_mode = 0;
void orc_set_x86_cvt(int on) { orc_x86_cvt_mode = on; }
void orc_set_libm_f64(int on) { orc_libm_f64_mode = on; }

_Static_assert(sizeof(orc_biquad_f32) == 68, "Biquad (RP2350) is 68 bytes, config.h:418-431");
_Static_assert(offsetof(orc_biquad_f32, s1) == 20 && offsetof(orc_biquad_f32, sva1) == 28 &&
               offsetof(orc_biquad_f32, svm0) == 40 && offsetof(orc_biquad_f32, svic1eq) == 52 &&
               offsetof(orc_biquad_f32, svf_type) == 60 && offsetof(orc_biquad_f32, use_svf) == 64 &&
               offsetof(orc_biquad_f32, bypass) == 65, "Biquad (RP2350) offsets");
_Static_assert(sizeof(orc_biquad_q28) == 32 && offsetof(orc_biquad_q28, bypass) == 28,
               "Biquad (RP2040) is 32 bytes, dsp_process_rp2040.S:6-14");
_Static_assert(sizeof(orc_eq_param) == 16, "EqParamPacket");
_Static_assert(sizeof(orc_crosspoint) == 12 && sizeof(orc_output) == 20, "matrix records");
_Static_assert(sizeof(orc_xfeed_f32) == 28 && sizeof(orc_xfeed_q28) == 28, "CrossfeedState");
_Static_assert(sizeof(orc_lev_coeffs) == 36, "LevellerCoeffs");
_Static_assert(sizeof(orc_lev_state_f32) == 3864 && sizeof(orc_lev_state_q28) == 3864, "LevellerState");
_Static_assert(sizeof(orc_loud_f32) == 28 && sizeof(orc_loud_q28) == 24, "LoudnessCoeffs");

size_t orc_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(orc_biquad_f32);
    case 1: return sizeof(orc_biquad_q28);
    case 2: return sizeof(orc_chain_f32);
    case 3: return sizeof(orc_chain_q28);
    case 4: return sizeof(orc_lev_state_f32);
    case 5: return sizeof(orc_lev_state_q28);
    default: return 0;
    }
}

/* float -> int32 cast.  Firmware semantics: saturate, NaN -> 0 (ARM). */
int32_t orc_f2i_sat(float x)
{
    if (orc_x86_cvt_mode) {
        if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT32_MIN;   /* CVTTSS2SI "integer indefinite" */
        return (int32_t)x;
    }
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

/* dsp_pipeline.c:47-58.  The lo*lo partial product is dropped and every
 * intermediate wraps at 32 bits; `ah * bl` is evaluated in unsigned arithmetic
 * (int32 * uint32) and converted back, i.e. the low 32 bits of the product. */
int32_t orc_mul_q28(int32_t a, int32_t b)
{
    int32_t  ah = a >> 16;
    uint32_t al = (uint32_t)a & 0xFFFFu;
    int32_t  bh = b >> 16;
    uint32_t bl = (uint32_t)b & 0xFFFFu;
    int32_t high = (int32_t)((uint32_t)ah * (uint32_t)bh);
    int32_t mid1 = (int32_t)((uint32_t)ah * bl);
    int32_t mid2 = (int32_t)(al * (uint32_t)bh);
    int32_t mid  = (int32_t)((uint32_t)mid1 + (uint32_t)mid2);
    return (int32_t)(((uint32_t)high << 4) + (uint32_t)(mid >> 12));
}

/* config.h:556-567 */
int32_t orc_mul_q15(int32_t sample, int32_t gain)
{
    int32_t  sh = sample >> 16;
    uint32_t sl = (uint16_t)sample;
    int32_t  gh = gain >> 16;
    uint32_t gl = (uint16_t)gain;
    int32_t  hh = (int32_t)((uint32_t)sh * (uint32_t)gh);
    int32_t  mid = (int32_t)((uint32_t)sh * gl + sl * (uint32_t)gh);
    uint32_t ll = sl * gl;
    return (int32_t)(((uint32_t)hh << 17) + ((uint32_t)mid << 1) + (ll >> 15));
}

/* dsp_process_rp2040.S:225-394.  Register roles in the assembly: r9=s1,
 * r10=s2, r6/r7 = high/low half of the operand being multiplied, r12 = b2*x.
 * Each inlined multiply (:273-283 etc.) is orc_mul_q28(coeff, operand). */
static void q28_eq_block(orc_biquad_q28 *bq, int32_t *samples, uint32_t count, uint32_t nbands)
{
    for (uint32_t band = 0; band < nbands; band++, bq++) {
        if (bq->bypass) continue;                                   /* :246-248 */
        uint32_t s1 = (uint32_t)bq->s1, s2 = (uint32_t)bq->s2;      /* :251-254 */
        for (uint32_t i = 0; i < count; i++) {                      /* :263-365 */
            int32_t x = samples[i];
            uint32_t y  = (uint32_t)orc_mul_q28(bq->b0, x) + s1;    /* :273-285 */
            uint32_t t1 = (uint32_t)orc_mul_q28(bq->b1, x);         /* :288-298 */
            uint32_t t3 = (uint32_t)orc_mul_q28(bq->b2, x);         /* :301-312 */
            uint32_t t2 = (uint32_t)orc_mul_q28(bq->a1, (int32_t)y);/* :319-329 */
            s1 = (t1 - t2) + s2;                                    /* :332-335 */
            uint32_t t4 = (uint32_t)orc_mul_q28(bq->a2, (int32_t)y);/* :338-348 */
            s2 = t3 - t4;                                           /* :351-353 */
            samples[i] = (int32_t)y;                                /* :356-357 */
        }
        bq->s1 = (int32_t)s1;                                       /* :370-373 */
        bq->s2 = (int32_t)s2;
    }
}

void orc_q28_eq_block(orc_biquad_q28 *bq, int32_t *samples, uint32_t count, uint32_t nbands)
{
    q28_eq_block(bq, samples, count, nbands);
}

void orc_q28_eq_many(orc_biquad_q28 *bq, int32_t *samples, uint32_t C, uint32_t T, uint32_t nbands, uint32_t packet)
{
    if (packet == 0) packet = T;
    for (uint32_t c = 0; c < C; c++) {
        orc_biquad_q28 *b = bq + (size_t)c * ORC_MAX_BANDS;
        int32_t *s = samples + (size_t)c * T;
        for (uint32_t t0 = 0; t0 < T; t0 += packet) {
            uint32_t n = (T - t0 < packet) ? (T - t0) : packet;
            q28_eq_block(b, s + t0, n, nbands);
        }
    }
}

/* crossfeed.c:161-180 */
static inline void q28_xfeed1(orc_xfeed_q28 *st, int32_t *left, int32_t *right)
{
    int32_t in_L = *left, in_R = *right;
    int32_t lp_L = orc_mul_q28(st->lp_a0, in_L) + orc_mul_q28(st->lp_b1, st->lp_state_L);   /* :166 */
    int32_t lp_R = orc_mul_q28(st->lp_a0, in_R) + orc_mul_q28(st->lp_b1, st->lp_state_R);   /* :167 */
    st->lp_state_L = lp_L;
    st->lp_state_R = lp_R;
    int32_t ap_L = orc_mul_q28(st->ap_a, lp_L) + st->ap_state_L;                            /* :172 */
    st->ap_state_L = lp_L - orc_mul_q28(st->ap_a, ap_L);                                    /* :173 */
    int32_t ap_R = orc_mul_q28(st->ap_a, lp_R) + st->ap_state_R;
    st->ap_state_R = lp_R - orc_mul_q28(st->ap_a, ap_R);
    *left  = (in_L - lp_L) + ap_R;                                                          /* :178 */
    *right = (in_R - lp_R) + ap_L;
}

void orc_q28_xfeed(orc_xfeed_q28 *st, int32_t *l, int32_t *r, uint32_t count)
{
    for (uint32_t i = 0; i < count; i++) q28_xfeed1(st, &l[i], &r[i]);
}

/* leveller.c:124-139 (float helper shared with the RP2040 build) */
static inline float q28_gain_computer(float x_db, float threshold, float ratio, float knee_width)
{
    float half_knee = knee_width * 0.5f;
    if (x_db > (threshold + half_knee)) return 0.0f;
    if (x_db >= (threshold - half_knee)) {
        float d = threshold + half_knee - x_db;
        return (1.0f - 1.0f / ratio) * d * d / (2.0f * knee_width);
    }
    return (threshold - x_db) * (1.0f - 1.0f / ratio);
}

/* leveller.c:275-389 */
static void q28_leveller(orc_lev_state_q28 *st, const orc_lev_coeffs *c, int lookahead,
                         int32_t *buf_l, int32_t *buf_r, uint32_t count)
{
    if (count == 0) return;
    const int32_t unity = 1 << 28;
    int32_t a_rms = orc_f2i_sat(c->alpha_rms * (float)(1 << 28));        /* :286 */
    int32_t one_minus = unity - a_rms;                                   /* :287 */
    int32_t env_l = st->env_sq_l, env_r = st->env_sq_r;
    for (uint32_t i = 0; i < count; i++) {                               /* :292-299 */
        int32_t sl = buf_l[i], sr = buf_r[i];
        int32_t sq_l = orc_mul_q28(sl, sl), sq_r = orc_mul_q28(sr, sr);
        env_l = orc_mul_q28(a_rms, env_l) + orc_mul_q28(one_minus, sq_l);
        env_r = orc_mul_q28(a_rms, env_r) + orc_mul_q28(one_minus, sq_r);
    }
    st->env_sq_l = env_l;
    st->env_sq_r = env_r;

    const float inv_q28 = 1.0f / (float)(1 << 28);                       /* :307 */
    float env_l_f = (float)env_l * inv_q28, env_r_f = (float)env_r * inv_q28;
    float rms_sq = (env_l_f > env_r_f) ? env_l_f : env_r_f;
    float rms_db = 10.0f * orc_log10f(rms_sq + 1e-30f);                  /* :311 */
    float gc_db;
    if (rms_db < c->gate_threshold_db) {
        gc_db = 0.0f;
    } else {
        gc_db = q28_gain_computer(rms_db, c->threshold_db, c->ratio, c->knee_width_db);
        gc_db += c->makeup_db;
        if (gc_db > c->max_gain_db) gc_db = c->max_gain_db;
    }
    float alpha_sample = (gc_db < st->gain_smooth_db) ? c->alpha_attack : c->alpha_release;
    float alpha = orc_powf(alpha_sample, (float)count);                  /* :327 */
    st->gain_smooth_db = alpha * st->gain_smooth_db + (1.0f - alpha) * gc_db;    /* :328-329 (RP2040: soft float, never fused) */
    float gain_linear = orc_powf(10.0f, st->gain_smooth_db / 20.0f);     /* :332 */
    st->gain_prev_q28 = st->gain_q28;
    st->gain_q28 = orc_f2i_sat(gain_linear * (float)(1 << 28));          /* :334 */

    int32_t g_prev = st->gain_prev_q28, g_cur = st->gain_q28;
    const float ceil_ = 0.70795f;
    uint32_t la_idx = st->la_write_idx;
    for (uint32_t i = 0; i < count; i++) {                               /* :347-386 */
        int32_t gain;
        if (count == 1) gain = g_cur;
        else gain = g_prev + (int32_t)(((int64_t)(g_cur - g_prev) * i) / (int32_t)(count - 1));   /* :352 */
        int32_t out_l, out_r;
        if (lookahead) {
            out_l = st->lookahead_buf[0][la_idx];
            out_r = st->lookahead_buf[1][la_idx];
            st->lookahead_buf[0][la_idx] = buf_l[i];
            st->lookahead_buf[1][la_idx] = buf_r[i];
            la_idx++;
            if (la_idx >= ORC_LA_SAMPLES) la_idx = 0;
        } else {
            out_l = buf_l[i];
            out_r = buf_r[i];
        }
        if (gain > unity) {                                              /* :370-379 */
            float peak = fabsf((float)out_l * inv_q28);
            float pr = fabsf((float)out_r * inv_q28);
            if (pr > peak) peak = pr;
            if (peak > 0.0f) {
                float max_g_f = ceil_ / peak;
                int32_t max_g = orc_f2i_sat(max_g_f * (float)unity);
                if (max_g < gain) gain = (max_g > unity) ? max_g : unity;
            }
        }
        buf_l[i] = orc_mul_q28(out_l, gain);
        buf_r[i] = orc_mul_q28(out_r, gain);
    }
    st->la_write_idx = la_idx;
}

void orc_q28_leveller(orc_lev_state_q28 *st, const orc_lev_coeffs *c, int lookahead, int32_t *l, int32_t *r, uint32_t count)
{
    unsigned csr = orc_ftz_enter();
    q28_leveller(st, c, lookahead, l, r, count);
    orc_ftz_leave(csr);
}

/* ---- ΔΣ PDM core ---------------------------------------------------------- */

void orc_pdm_reset(orc_pdm_state *st)
{
    memset(st, 0, sizeof(*st));
    st->rng = 123456789u;                                                /* pdm_generator.c:62 */
}

/* pdm_generator.c:351-397, steady-state branch (hardware running, a sample is
 * available, no fade-out).  One Q28 input sample -> 8 words = 256 PDM bits,
 * MSB first. */
void orc_pdm_modulate(orc_pdm_state *st, int32_t sample_q28, uint32_t out[8])
{
    int32_t pcm = sample_q28 >> 14;                                      /* :352 */
    if (pcm > 29500) pcm = 29500;                                        /* :353-354, config.h:64 */
    if (pcm < -29500) pcm = -29500;
    if (st->fade_in_pos < 1024u) {                                       /* :357-360 */
        pcm = (pcm * (int32_t)st->fade_in_pos) >> 10;
        st->fade_in_pos++;
    }
    int32_t target = pcm + 32768;                                        /* :363 */
    int32_t err1 = st->err1, err2 = st->err2;
    for (int chunk = 0; chunk < 8; chunk++) {                            /* :367 */
        uint32_t r = st->rng;                                            /* :63-68 xorshift32 */
        r ^= r << 13; r ^= r >> 17; r ^= r << 5;
        st->rng = r;
        int32_t raw = (int32_t)(r & 0x1FFu) - 255;                       /* :368 */
        /* noise_shaped_dither(), :89-108 */
        int32_t qerr = err2 >> 8;
        st->err_acc = ((st->err_acc * 248) >> 8) + (qerr >> 6);          /* :92 */
        int32_t input = raw - st->err_acc;                               /* :95 */
        int32_t dither = (15778 * input + (-31556) * st->x1 + 15778 * st->x2
                          + 31531 * st->y1 - 15580 * st->y2) >> 14;      /* :98-99 */
        st->x2 = st->x1; st->x1 = input;
        st->y2 = st->y1; st->y1 = dither;
        uint32_t word = 0;
        for (int k = 0; k < 32; k++) {                                   /* :372-378 */
            int bit = (err2 + dither) >= 0;
            int32_t fb = bit ? 65535 : 0;
            if (bit) word |= 1u << (31 - k);
            err1 += target - fb;
            err2 += err1 - fb;
        }
        out[chunk] = word;                                               /* :380 */
    }
    err1 -= err1 >> 16;                                                  /* :396-397 */
    err2 -= err2 >> 16;
    st->err1 = err1;
    st->err2 = err2;
}

/* ---- whole packet, RP2040 pipeline (usb_audio.c:968-1283, single-core :1191-1276) */
uint32_t orc_q28_chain_packet(orc_chain_q28 *in, const uint8_t *data, uint32_t data_len, uint32_t bit_depth,
                              int32_t *spdif_out, uint32_t spdif_stride, uint32_t *pdm_out)
{
    unsigned csr = orc_ftz_enter();
    static __thread int32_t buf_l[ORC_PKT_MAX], buf_r[ORC_PKT_MAX], buf_out[ORC_MAX_OUT][ORC_PKT_MAX];
    const uint32_t bytes_per_frame = (bit_depth == 24) ? 6 : 4;
    uint32_t n = data_len / bytes_per_frame;
    if (n > ORC_PKT_MAX) n = ORC_PKT_MAX;
    const uint32_t O = in->n_out;
    const uint32_t mask = in->max_delay - 1;

    if (in->mute_env_on)                                                 /* :532 */
        in->preset_mute_gain = orc_mute_envelope(&in->preset_loading, &in->preset_mute_counter, &in->preset_mute_smooth_gain,
                                                 data_len / ((bit_depth == 24) ? 6u : 4u), in->sample_rate_hz);
    int32_t vol_mul = in->host_mute ? 0 : (int32_t)in->host_vol_mul;                    /* :975 */
    int32_t pmg = (int32_t)(in->preset_mute_gain * 32768.0f + 0.5f);                    /* :976 */
    if (pmg < 0) pmg = 0;
    if (pmg > 32768) pmg = 32768;
    vol_mul = orc_mul_q15(vol_mul, pmg);                                                /* :979 */
    int32_t vol_mul_master = orc_mul_q15(vol_mul, in->master_volume_q15);               /* :980 */
    int32_t preamp_l = in->preamp_q28[0], preamp_r = in->preamp_q28[1];

    /* PASS 1 (:997-1015) */
    if (bit_depth == 24) {
        const uint8_t *p = data;
        for (uint32_t i = 0; i < n; i++, p += 6) {
            int32_t l = (int32_t)((uint32_t)p[2] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[0] << 8) >> 2;
            int32_t r = (int32_t)((uint32_t)p[5] << 24 | (uint32_t)p[4] << 16 | (uint32_t)p[3] << 8) >> 2;
            buf_l[i] = orc_mul_q28(l, preamp_l);
            buf_r[i] = orc_mul_q28(r, preamp_r);
        }
    } else {
        for (uint32_t i = 0; i < n; i++) {
            int16_t l = (int16_t)((uint16_t)data[i * 4 + 0] | (uint16_t)data[i * 4 + 1] << 8);
            int16_t r = (int16_t)((uint16_t)data[i * 4 + 2] | (uint16_t)data[i * 4 + 3] << 8);
            buf_l[i] = orc_mul_q28((int32_t)((uint32_t)(int32_t)l << 14), preamp_l);
            buf_r[i] = orc_mul_q28((int32_t)((uint32_t)(int32_t)r << 14), preamp_r);
        }
    }
    /* loudness (:1018-1047) */
    if (in->loudness_on) {
        for (uint32_t i = 0; i < n; i++) {
            int32_t raw[2] = { buf_l[i], buf_r[i] };
            for (int side = 0; side < 2; side++) {
                int32_t x = raw[side];
                for (int j = 0; j < 2; j++) {
                    const orc_loud_q28 *lc = &in->loud[j];
                    if (lc->bypass) continue;
                    orc_biquad_q28 *bq = &in->loud_state[side][j];
                    int32_t result = orc_mul_q28(lc->b0, x) + bq->s1;
                    bq->s1 = orc_mul_q28(lc->b1, x) - orc_mul_q28(lc->a1, result) + bq->s2;
                    bq->s2 = orc_mul_q28(lc->b2, x) - orc_mul_q28(lc->a2, result);
                    x = result;
                }
                raw[side] = x;
            }
            buf_l[i] = raw[0];
            buf_r[i] = raw[1];
        }
    }
    /* PASS 2 (:1050-1055), 2.5 (:1058-1062) */
    const int is_bypassed = in->bypass_master_eq;
    if (!is_bypassed) {
        if (!in->channel_bypassed[0]) q28_eq_block(in->filters[0], buf_l, n, in->n_bands);
        if (!in->channel_bypassed[1]) q28_eq_block(in->filters[1], buf_r, n, in->n_bands);
    }
    if (in->leveller_on) q28_leveller(&in->levs, &in->levc, in->lev_lookahead, buf_l, buf_r, n);
    /* PASS 3 (:1065-1073) */
    int32_t peak_ml = 0, peak_mr = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t ml = buf_l[i], mr = buf_r[i];
        if (abs(ml) > peak_ml) peak_ml = abs(ml);
        if (abs(mr) > peak_mr) peak_mr = abs(mr);
        if (in->crossfeed_on) {
            q28_xfeed1(&in->xfeed, &ml, &mr);
            buf_l[i] = ml; buf_r[i] = mr;
        }
    }
    /* PASS 4 (:1076-1100) — crosspoint gains quantised to Q15 every packet */
    for (uint32_t o = 0; o < O; o++) {
        int32_t *dst = buf_out[o];
        if (!in->out[o].enabled) { memset(dst, 0, n * sizeof(int32_t)); continue; }
        const orc_crosspoint *xl = &in->xp[0][o], *xr = &in->xp[1][o];
        int32_t gl = xl->enabled ? orc_f2i_sat((xl->phase_invert ? -xl->gain_linear : xl->gain_linear) * 32768.0f) : 0;
        int32_t gr = xr->enabled ? orc_f2i_sat((xr->phase_invert ? -xr->gain_linear : xr->gain_linear) * 32768.0f) : 0;
        if (gl != 0 && gr != 0) for (uint32_t i = 0; i < n; i++) dst[i] = orc_mul_q15(buf_l[i], gl) + orc_mul_q15(buf_r[i], gr);
        else if (gl != 0)       for (uint32_t i = 0; i < n; i++) dst[i] = orc_mul_q15(buf_l[i], gl);
        else if (gr != 0)       for (uint32_t i = 0; i < n; i++) dst[i] = orc_mul_q15(buf_r[i], gr);
        else                    memset(dst, 0, n * sizeof(int32_t));
    }
    /* PASS 5 (:1196-1213) — RP2040 also gates output EQ on bypass_master_eq (quirk 3) */
    for (uint32_t o = 0; o < O; o++) {
        if (!in->out[o].enabled) continue;
        if (!in->out[o].mute) {
            uint32_t eq_ch = 2 + o;
            if (!is_bypassed && !in->channel_bypassed[eq_ch]) q28_eq_block(in->filters[eq_ch], buf_out[o], n, in->n_bands);
        }
        int32_t gain = in->out[o].mute ? 0 : orc_f2i_sat(in->out[o].gain_linear * (float)vol_mul_master);   /* :1204-1205 */
        if (gain == 0) memset(buf_out[o], 0, n * sizeof(int32_t));
        else for (uint32_t i = 0; i < n; i++) buf_out[o][i] = orc_mul_q15(buf_out[o][i], gain);
    }
    /* PASS 6 (:1216-1230) */
    if (in->any_delay_active) {
        for (uint32_t o = 0; o < O; o++) {
            int32_t dly = in->delay_samples[o];
            if (dly <= 0) continue;
            int32_t *dst = buf_out[o], *dline = in->delay_lines[o];
            uint32_t w = in->delay_widx;
            for (uint32_t i = 0; i < n; i++) {
                dline[w] = dst[i];
                dst[i] = dline[(w - (uint32_t)dly) & mask];
                w = (w + 1) & mask;
            }
        }
        in->delay_widx = (in->delay_widx + n) & mask;
    }
    /* PASS 7 (:1233-1257) */
    const uint32_t n_spdif = O - 1;
    const int32_t clip_thresh = (1 << 28) + 268;                         /* config.h:54 */
    for (uint32_t o = 0; o < n_spdif; o++) {
        int32_t peak = 0;
        for (uint32_t i = 0; i < n; i++) { int32_t a = abs(buf_out[o][i]); if (a > peak) peak = a; }
        in->peaks[2 + o] = (uint16_t)(peak >> 13);
        if (peak > clip_thresh) in->clip_flags |= (uint16_t)(1u << (2 + o));
    }
    for (uint32_t pair = 0; pair < n_spdif / 2; pair++) {
        uint32_t lc = pair * 2, rc = pair * 2 + 1;
        int32_t *op = spdif_out + (size_t)pair * spdif_stride;
        if (!in->out[lc].enabled && !in->out[rc].enabled) { memset(op, 0, (size_t)n * 8); continue; }
        for (uint32_t i = 0; i < n; i++) {
            int32_t a = (buf_out[lc][i] + 32) >> 6, b = (buf_out[rc][i] + 32) >> 6;      /* :1254-1255 */
            op[i * 2]     = a > 0x7FFFFF ? 0x7FFFFF : (a < -0x800000 ? -0x800000 : a);   /* config.h:547-551 */
            op[i * 2 + 1] = b > 0x7FFFFF ? 0x7FFFFF : (b < -0x800000 ? -0x800000 : b);
        }
    }
    /* PDM sub (:1261-1274) */
    const uint32_t sub = O - 1;
    if (in->out[sub].enabled) {
        int32_t peak = 0;
        for (uint32_t i = 0; i < n; i++) { int32_t a = abs(buf_out[sub][i]); if (a > peak) peak = a; }
        in->peaks[2 + sub] = (uint16_t)(peak >> 13);
        if (peak > clip_thresh) in->clip_flags |= (uint16_t)(1u << (2 + sub));
        for (uint32_t i = 0; i < n; i++) orc_pdm_modulate(&in->pdm, buf_out[sub][i], pdm_out + (size_t)i * 8);
    } else {
        in->peaks[2 + sub] = 0;
    }
    in->peaks[0] = (uint16_t)(peak_ml >> 13);                            /* :1279-1282 */
    in->peaks[1] = (uint16_t)(peak_mr >> 13);
    if (peak_ml > clip_thresh) in->clip_flags |= 1u;
    if (peak_mr > clip_thresh) in->clip_flags |= 2u;
    orc_ftz_leave(csr);
    return n;
}
