// dynamics.cuh — parameter ingest on the device, second half: the crossfeed, leveller and
// loudness coefficient generators plus audio_set_volume(), as device functions shared by the two chain engines.
//
// Reference: crossfeed_compute_coefficients() crossfeed.c:35-127, leveller_compute_coefficients() leveller.c:42-89
// (time constants :23-27, retention :37-40), loudness_recompute_table() loudness.c:169-217 (ISO 226 rows :20-28,
// equal-loudness level :37-50, compensation :54-78, shelf design :85-163), audio_set_volume() usb_audio.c:410-440.
//
// Same rules as coeff.cu: every float operation is the reference's, rounded on its own (-fmad=false, IEEE division and
// square root spelled as intrinsics so that nvcc cannot turn `x / constant` into a multiplication); the libm calls
// (powf, expf, logf, log10f, tanf, sinf, cosf) follow the libm policy of DESIGN.md §6 - evaluated in double, rounded once.
#pragma once
#include <cstdint>
#include "dspi_b200.h"

namespace dspi {
namespace dyn {

constexpr float kPi = 3.1415926535f;                          // the reference's literal

__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float pow_f(float a, float b) { return (float)pow((double)a, (double)b); }
__device__ __forceinline__ float exp_f(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_f(float x) { return (float)log((double)x); }
__device__ __forceinline__ float log10_f(float x) { return (float)log10((double)x); }
__device__ __forceinline__ float tan_f(float x) { return (float)tan((double)x); }
__device__ __forceinline__ float sin_f(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos_f(float x) { return (float)cos((double)x); }
__device__ __forceinline__ int32_t f2i_sat(float x) { return __float2int_rz(x); }       // cvt.rzi saturates, NaN -> 0 (ARM VCVT)

// crossfeed.c:35-127.  Returns false when the stage is disabled (state record all zero, crossfeed_init()).
__device__ inline bool crossfeed_coeffs(const dspi_crossfeed_config &cfg, float fs, float &lp_a0, float &lp_b1, float &ap_a)
{
    lp_a0 = lp_b1 = ap_a = 0.0f;
    if (!cfg.enabled || fs < 1.0f) return false;
    float fc, feed_db;
    if (cfg.preset < 3) {                                     // :25-29 presets
        fc = cfg.preset == 2 ? 650.0f : 700.0f;
        feed_db = cfg.preset == 0 ? 4.5f : (cfg.preset == 1 ? 6.0f : 9.5f);
    } else {                                                  // :47-52 custom, clamped
        fc = cfg.custom_fc;
        feed_db = cfg.custom_feed_db;
        if (fc < 500.0f) fc = 500.0f;
        if (fc > 2000.0f) fc = 2000.0f;
        if (feed_db < 0.0f) feed_db = 0.0f;
        if (feed_db > 15.0f) feed_db = 15.0f;
    }
    const float level_ratio = pow_f(10.0f, fdiv(feed_db, 20.0f));     // :67
    const float G = fdiv(1.0f, 1.0f + level_ratio);
    const float x = exp_f(fdiv(-2.0f * kPi * fc, fs));                 // :75
    lp_a0 = G * (1.0f - x);
    lp_b1 = x;
    ap_a = 1.0f;
    if (cfg.itd_enabled) {                                    // :97-109
        const float lp_delay_sec = fdiv(x, (1.0f - x) * fs);
        const float remaining = 0.000220f - lp_delay_sec;
        if (remaining > 0.0f) {
            const float D = remaining * fs;
            ap_a = fdiv(1.0f - D, 1.0f + D);
        }
    }
    return true;
}

// leveller.c:37-40
__device__ inline float retention(float fs, float seconds)
{
    if (seconds <= 0.0f || fs <= 0.0f) return 0.0f;
    return exp_f(fdiv(-log_f(10.0f), fs * seconds));
}

// leveller.c:42-89 -> LevellerCoeffs in field order (alpha_rms, alpha_attack, alpha_release, threshold_db, ratio,
// knee_width_db, makeup_db, gate_threshold_db, max_gain_db)
__device__ inline void leveller_coeffs(const dspi_leveller_config &cfg, float fs, float (&out)[9])
{
    if (fs < 1.0f) fs = 48000.0f;
    const unsigned spd = cfg.speed >= 3 ? 1u : cfg.speed;    // {attack, release, rms window} seconds, :23-27
    const float attack = spd == 0 ? 0.100f : (spd == 1 ? 0.050f : 0.020f);
    const float release = spd == 0 ? 2.000f : (spd == 1 ? 1.000f : 0.500f);
    const float window = spd == 0 ? 0.400f : (spd == 1 ? 0.200f : 0.100f);
    out[0] = retention(fs, window);
    out[1] = retention(fs, attack);
    out[2] = retention(fs, release);
    out[3] = -20.0f;                                          // leveller.h:51-52
    float amount = cfg.amount;
    if (amount < 0.0f) amount = 0.0f;
    if (amount > 100.0f) amount = 100.0f;
    out[4] = 1.0f + fdiv(amount, 100.0f) * 19.0f;             // :76-77
    out[5] = 6.0f;
    out[6] = 0.0f;
    float gate = cfg.gate_threshold_db;
    if (gate < -96.0f) gate = -96.0f;
    if (gate > 0.0f) gate = 0.0f;
    out[7] = gate;
    float max_g = cfg.max_gain_db;
    if (max_g < 0.0f) max_g = 0.0f;
    if (max_g > 35.0f) max_g = 35.0f;
    out[8] = max_g;
}

// loudness.c:37-50
__device__ inline float iso226_spl(float Tf, float af, float Lu, float phon)
{
    const float B = 0.4f * pow_f(10.0f, fdiv(Tf + Lu, 10.0f) - 9.0f);
    const float threshold = pow_f(B, af);
    float Af = 4.47e-3f * (pow_f(10.0f, 0.025f * phon) - 1.15f) + threshold;
    if (Af < 1e-10f) Af = 1e-10f;
    return fdiv(10.0f, af) * log10_f(Af) - Lu + 94.0f;
}

// loudness.c:54-78
__device__ inline float compensation_db(float Tf, float af, float Lu, float ref_spl, float phon, float intensity_pct)
{
    if (phon >= ref_spl) return 0.0f;
    const float at_ref = iso226_spl(Tf, af, Lu, ref_spl);
    const float at_eff = iso226_spl(Tf, af, Lu, phon);
    const float flat_change = phon - ref_spl;
    const float freq_change = at_eff - at_ref;
    float comp = freq_change - flat_change;
    comp *= fdiv(intensity_pct, 100.0f);
    return comp;
}

// loudness.c:169-217 for ONE volume step: the two shelf gains of table row `step` (0..60)
__device__ inline void loudness_row_gains(int step, float ref_spl, float intensity_pct, float &low_db, float &high_db)
{
    if (ref_spl < 40.0f) ref_spl = 40.0f;
    if (ref_spl > 100.0f) ref_spl = 100.0f;
    float phon = ref_spl + (float)(step - 60);                // :186-191
    if (phon < 20.0f) phon = 20.0f;
    if (phon > ref_spl) phon = ref_spl;
    low_db = compensation_db(44.0f, 0.432f, 80.4f, ref_spl, phon, intensity_pct);      // 50 Hz row of ISO 226
    high_db = compensation_db(13.9f, 0.301f, 17.8f, ref_spl, phon, intensity_pct);     // 10 kHz row
}

// loudness.c:85-130, float branch: c = {sva1, sva2, sva3, svm0, svm1, svm2}
__device__ inline void shelf_svf(float freq, float Q, float gain_db, bool high, float fs, float (&c)[6], bool &bypass)
{
    for (int k = 0; k < 6; k++) c[k] = 0.0f;
    bypass = fabsf(gain_db) < 0.01f;
    if (bypass) return;
    const float A = pow_f(10.0f, fdiv(gain_db, 40.0f));
    float g = tan_f(fdiv(kPi * freq, fs));
    const float rootA = __fsqrt_rn(A);
    g = high ? g * rootA : fdiv(g, rootA);
    const float k = fdiv(1.0f, Q);
    c[0] = fdiv(1.0f, 1.0f + g * (g + k));
    c[1] = g * c[0];
    c[2] = g * c[1];
    if (high) { c[3] = A * A; c[4] = k * (1.0f - A) * A; c[5] = 1.0f - A * A; }
    else { c[3] = 1.0f; c[4] = k * (A - 1.0f); c[5] = A * A - 1.0f; }
}

// loudness.c:131-163, Q28 branch (RBJ shelf, truncating Q28 store): c = {b0, b1, b2, a1, a2}
__device__ inline void shelf_q28(float freq, float Q, float gain_db, bool high, float fs, int32_t (&c)[5], bool &bypass)
{
    bypass = fabsf(gain_db) < 0.01f;
    if (bypass) { c[0] = 1 << 28; c[1] = c[2] = c[3] = c[4] = 0; return; }
    const float A = pow_f(10.0f, fdiv(gain_db, 40.0f));
    const float omega = fdiv(2.0f * kPi * freq, fs);
    const float sn = sin_f(omega), cs = cos_f(omega);
    const float alpha = fdiv(sn, 2.0f * Q);
    const float sA = __fsqrt_rn(A);
    float a0, a1, a2, b0, b1, b2;
    if (high) {
        b0 = A * ((A + 1) + (A - 1) * cs + 2 * sA * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * cs);
        b2 = A * ((A + 1) + (A - 1) * cs - 2 * sA * alpha);
        a0 = (A + 1) - (A - 1) * cs + 2 * sA * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * cs);
        a2 = (A + 1) - (A - 1) * cs - 2 * sA * alpha;
    } else {
        b0 = A * ((A + 1) - (A - 1) * cs + 2 * sA * alpha);
        b1 = 2 * A * ((A - 1) - (A + 1) * cs);
        b2 = A * ((A + 1) - (A - 1) * cs - 2 * sA * alpha);
        a0 = (A + 1) + (A - 1) * cs + 2 * sA * alpha;
        a1 = -2 * ((A - 1) + (A + 1) * cs);
        a2 = (A + 1) + (A - 1) * cs - 2 * sA * alpha;
    }
    const float scale = 268435456.0f;
    c[0] = f2i_sat(fdiv(b0, a0) * scale);
    c[1] = f2i_sat(fdiv(b1, a0) * scale);
    c[2] = f2i_sat(fdiv(b2, a0) * scale);
    c[3] = f2i_sat(fdiv(a1, a0) * scale);
    c[4] = f2i_sat(fdiv(a2, a0) * scale);
}

// audio_set_volume(), usb_audio.c:410-440: UAC1 volume (1/256 dB) -> Q15 multiplier stored in an int16 (0 dB wraps to
// -32768) and the loudness table row
__device__ inline int16_t host_volume(int16_t volume_8_8, uint32_t &row)
{
    const uint16_t q15[61] = {
        0x0000, 0x0025, 0x0029, 0x002e, 0x0034, 0x003a, 0x0041, 0x0049, 0x0052, 0x005c, 0x0068, 0x0074, 0x0082, 0x0092, 0x00a4, 0x00b8,
        0x00cf, 0x00e8, 0x0104, 0x0124, 0x0148, 0x0170, 0x019d, 0x01cf, 0x0207, 0x0247, 0x028e, 0x02de, 0x0337, 0x039c, 0x040c, 0x048b,
        0x0519, 0x05b8, 0x066a, 0x0733, 0x0814, 0x0910, 0x0a2b, 0x0b68, 0x0ccd, 0x0e5d, 0x101d, 0x1215, 0x1449, 0x16c3, 0x198a, 0x1ca8,
        0x2027, 0x2413, 0x287a, 0x2d6b, 0x32f5, 0x392d, 0x4027, 0x47fb, 0x50c3, 0x5a9e, 0x65ad, 0x7215, 0x8000 };
    int16_t v = (int16_t)(volume_8_8 + 60 * 256);
    if (v < 0) v = 0;
    if (v >= 61 * 256) v = 61 * 256 - 1;
    row = ((uint16_t)v) >> 8;
    return (int16_t)q15[row];
}

}  // namespace dyn
}  // namespace dspi
