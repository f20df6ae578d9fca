/*
 * dspi_oracle.h — CPU ORACLE for the DSPi per-sample DSP signal chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (dspi_b200/) never links,
 * imports or falls back to this code.
 *
 * What it is: a plain-C restatement of the reference firmware's arithmetic
 * (WeebLabs/DSPi, /root/reference/firmware/DSPi) for the hot path of
 * SURVEY.md §8(a).  Every function cites the reference file:line it follows.
 *
 * Pinning status: the reference ships NO golden vectors or tests for this path, so the
 * restatement is pinned against the reference's OWN sources compiled on the host (oracle/_ref/, see
 * oracle/Makefile), bit for bit:
 *   - dsp_pipeline.c, crossfeed.c, leveller.c, loudness.c unmodified (ref_shim.c)      tests/test_oracle_vs_ref.py
 *   - usb_audio.c unmodified: process_audio_packet(), the whole per-packet orchestrator, both platform builds,
 *     float strict + fused (ref_chain_shim.c, SDK platform layer in stubs_fw/)          tests/test_chain_vs_ref_cpu.py
 *   - pdm_generator.c unmodified: the delta-sigma loop entered through pdm_core1_entry() (ref_pdm_shim.c)
 *   - bulk_params.c, flash_storage.c, sample_encoding.h for the "next" rows
 * and against committed fixtures generated from those builds (tests/golden/).  What stays a restatement without a
 * compiled counterpart: the Thumb-1 block cascade of dsp_process_rp2040.S (no ARM toolchain; restated from the listing
 * and cross-checked against a loop around the reference's compiled fast_mul_q28) and the two ARMv8-M inline-asm blocks
 * of the 24-bit unpack (usb_audio.c:613-674; restated instruction by instruction in ref_chain_shim.c).
 *
 * Three arithmetic flavours:
 *   f32 strict  — every multiply/add rounded separately (host gcc, no FMA)
 *   f32 fused   — GCC's -ffp-contract=fast contraction pattern, written out
 *                 with explicit fmaf(); this is what arm-none-eabi-gcc emits
 *                 for the RP2350 (Cortex-M33 VFMA) build of the firmware
 *   q28         — RP2040 fixed point, 32-bit wrapping, bit-exact
 * All float entry points run with FTZ|DAZ set (firmware: FPSCR FZ+DN,
 * main.c:593-600) and restore MXCSR on return.
 */
#ifndef DSPI_ORACLE_H
#define DSPI_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BANDS      12   /* config.h:329 MAX_BANDS (storage stride)        */
#define ORC_MAX_OUT         9   /* config.h:321 NUM_OUTPUT_CHANNELS (RP2350)      */
#define ORC_MAX_EQ_CH      11   /* config.h:322 NUM_CHANNELS (RP2350)             */
#define ORC_PKT_MAX       192   /* usb_audio.c:273,588 buf_l/buf_r/buf_out length */
#define ORC_LA_SAMPLES    480   /* leveller.h:36 LEVELLER_LOOKAHEAD_SAMPLES       */
#define ORC_MAX_DELAY    4096   /* config.h:84  (RP2350; RP2040 uses 2048)        */
#define ORC_LOUD_STEPS     61   /* loudness.h:7                                   */

/* filter types, config.h:440-443 */
enum { ORC_FLAT = 0, ORC_PEAKING = 1, ORC_LOWSHELF = 2, ORC_HIGHSHELF = 3,
       ORC_LOWPASS = 4, ORC_HIGHPASS = 5 };

/* ---- layout-compatible records (sizes/offsets asserted in dspi_oracle.c and
 *      checked against the compiled reference in tests) --------------------- */

/* config.h:418-431 (RP2350 Biquad, 68 bytes) */
typedef struct {
    float b0, b1, b2, a1, a2;
    float s1, s2;
    float sva1, sva2, sva3;
    float svm0, svm1, svm2;
    float svic1eq, svic2eq;
    uint32_t svf_type;
    uint8_t use_svf;
    uint8_t bypass;
} orc_biquad_f32;

/* config.h:433-437 / dsp_process_rp2040.S:6-14 (RP2040 Biquad, 32 bytes) */
typedef struct {
    int32_t b0, b1, b2, a1, a2;
    int32_t s1, s2;
    uint8_t bypass;
} orc_biquad_q28;

/* config.h:445-453 EqParamPacket (packed, 16 bytes) */
typedef struct __attribute__((packed)) {
    uint8_t channel, band, type, reserved;
    float freq, Q, gain_db;
} orc_eq_param;

/* config.h:383-389 / :392-400 (packed, 12 / 20 bytes) */
typedef struct __attribute__((packed)) {
    uint8_t enabled, phase_invert, reserved[2];
    float gain_db, gain_linear;
} orc_crosspoint;
typedef struct __attribute__((packed)) {
    uint8_t enabled, mute, reserved[2];
    float gain_db, gain_linear, delay_ms;
    int32_t delay_samples;
} orc_output;

/* crossfeed.h:26-32 CrossfeedConfig, :46-59 CrossfeedState (28 bytes) */
typedef struct {
    uint8_t enabled, itd_enabled, preset;
    float custom_fc, custom_feed_db;
} orc_xfeed_cfg;
typedef struct { float lp_a0, lp_b1, lp_state_L, lp_state_R, ap_a, ap_state_L, ap_state_R; } orc_xfeed_f32;
typedef struct { int32_t lp_a0, lp_b1, lp_state_L, lp_state_R, ap_a, ap_state_L, ap_state_R; } orc_xfeed_q28;

/* leveller.h:59-66 LevellerConfig, :81-99 LevellerCoeffs (36 B), :107-136 LevellerState (3864 B) */
typedef struct {
    uint8_t enabled;
    float amount;
    uint8_t speed;
    float max_gain_db;
    uint8_t lookahead;
    float gate_threshold_db;
} orc_lev_cfg;
typedef struct {
    float alpha_rms, alpha_attack, alpha_release;
    float threshold_db, ratio, knee_width_db, makeup_db, gate_threshold_db, max_gain_db;
} orc_lev_coeffs;
typedef struct {
    float env_sq_l, env_sq_r, gain_smooth_db, gain_linear, gain_prev_linear;
    float lookahead_buf[2][ORC_LA_SAMPLES];
    uint32_t la_write_idx;
} orc_lev_state_f32;
typedef struct {
    int32_t env_sq_l, env_sq_r;
    float gain_smooth_db;
    int32_t gain_q28, gain_prev_q28;
    int32_t lookahead_buf[2][ORC_LA_SAMPLES];
    uint32_t la_write_idx;
} orc_lev_state_q28;

/* loudness.h:11-23 (28 / 24 bytes) */
typedef struct { float sva1, sva2, sva3, svm0, svm1, svm2; uint8_t bypass; } orc_loud_f32;
typedef struct { int32_t b0, b1, b2, a1, a2; uint8_t bypass; } orc_loud_q28;
typedef struct { float ic1eq, ic2eq; } orc_svf_state;

/* pdm_generator.c:83-87 noise_shaper_t + loop locals :205-217 */
typedef struct {
    int32_t err1, err2;
    int32_t x1, x2, y1, y2, err_acc;
    uint32_t rng;
    uint32_t fade_in_pos;
} orc_pdm_state;

/* ---- whole-instance records (oracle's own layout) ------------------------ */

typedef struct {
    uint32_t n_out;              /* 9 */
    uint32_t n_bands;            /* channel_band_counts[] value (10)            */
    uint32_t max_delay;          /* MAX_DELAY_SAMPLES, power of two             */
    uint8_t  bypass_master_eq, loudness_on, crossfeed_on, leveller_on;
    uint8_t  host_mute, any_delay_active, lev_lookahead, pad0;
    int16_t  host_vol_mul;       /* AudioState.vol_mul (int16 — quirk 1)        */
    int16_t  pad1;
    float    preset_mute_gain;   /* update_preset_mute_envelope() result        */
    float    master_volume_linear;
    float    preamp_linear[2];
    orc_crosspoint xp[2][ORC_MAX_OUT];
    orc_output     out[ORC_MAX_OUT];
    int32_t  delay_samples[ORC_MAX_OUT];
    uint8_t  channel_bypassed[ORC_MAX_EQ_CH];
    uint8_t  pad2;
    orc_biquad_f32 filters[ORC_MAX_EQ_CH][ORC_MAX_BANDS];
    orc_loud_f32   loud[2];
    orc_svf_state  loud_state[2][2];
    orc_xfeed_f32  xfeed;
    orc_lev_coeffs levc;
    orc_lev_state_f32 levs;
    float    delay_lines[ORC_MAX_OUT][ORC_MAX_DELAY];
    uint32_t delay_widx;
    orc_pdm_state pdm;
    uint16_t peaks[ORC_MAX_EQ_CH];
    uint16_t clip_flags;
    /* preset-mute envelope (usb_audio.c:457-498, flash_storage.c:255-256).  mute_env_on == 0: the packet uses
     * preset_mute_gain above as given (a caller-computed constant); != 0: the envelope below runs once per
     * packet exactly as update_preset_mute_envelope() and preset_mute_gain is overwritten with its result. */
    uint8_t  mute_env_on, preset_loading, pad3[2];
    uint32_t preset_mute_counter;
    float    preset_mute_smooth_gain;
    uint32_t sample_rate_hz;
} orc_chain_f32;

typedef struct {
    uint32_t n_out;              /* 5 */
    uint32_t n_bands;
    uint32_t max_delay;          /* 2048 */
    uint8_t  bypass_master_eq, loudness_on, crossfeed_on, leveller_on;
    uint8_t  host_mute, any_delay_active, lev_lookahead, pad0;
    int16_t  host_vol_mul;
    int16_t  pad1;
    float    preset_mute_gain;
    int32_t  master_volume_q15;
    int32_t  preamp_q28[2];
    orc_crosspoint xp[2][ORC_MAX_OUT];
    orc_output     out[ORC_MAX_OUT];
    int32_t  delay_samples[ORC_MAX_OUT];
    uint8_t  channel_bypassed[ORC_MAX_EQ_CH];
    uint8_t  pad2;
    orc_biquad_q28 filters[ORC_MAX_EQ_CH][ORC_MAX_BANDS];
    orc_loud_q28   loud[2];
    orc_biquad_q28 loud_state[2][2];     /* only s1/s2 used (usb_audio.c:182) */
    orc_xfeed_q28  xfeed;
    orc_lev_coeffs levc;
    orc_lev_state_q28 levs;
    int32_t  delay_lines[ORC_MAX_OUT][ORC_MAX_DELAY];
    uint32_t delay_widx;
    orc_pdm_state pdm;
    uint16_t peaks[ORC_MAX_EQ_CH];
    uint16_t clip_flags;
    /* preset-mute envelope (usb_audio.c:457-498, flash_storage.c:255-256).  mute_env_on == 0: the packet uses
     * preset_mute_gain above as given (a caller-computed constant); != 0: the envelope below runs once per
     * packet exactly as update_preset_mute_envelope() and preset_mute_gain is overwritten with its result. */
    uint8_t  mute_env_on, preset_loading, pad3[2];
    uint32_t preset_mute_counter;
    float    preset_mute_smooth_gain;
    uint32_t sample_rate_hz;
} orc_chain_q28;

/* ---- control -------------------------------------------------------------- */

/* float->int32 conversion semantics for casts whose operand can leave the
 * int32 range.  0 (default) = saturating, NaN->0: ARM VCVT / __aeabi_f2iz and
 * CUDA cvt.rzi.s32.f32, i.e. what the firmware does.  1 = x86 CVTTSS2SI
 * (0x80000000 on overflow/NaN): only to pin the restatement against the
 * x86-compiled reference objects. */
void orc_set_x86_cvt(int on);
/* libm flavour for the leveller's per-block log10f/powf: 0 = glibc float
 * routines (matches oracle/_ref), 1 = double-precision evaluation rounded to
 * float (what the CUDA path computes; see DESIGN.md "libm policy"). */
void orc_set_libm_f64(int on);
size_t orc_sizeof(int which);   /* 0 biquad_f32, 1 biquad_q28, 2 chain_f32, 3 chain_q28, 4 lev_state_f32, 5 lev_state_q28 */

/* ---- arithmetic primitives ------------------------------------------------ */
int32_t orc_mul_q28(int32_t a, int32_t b);    /* dsp_pipeline.c:47-58  */
int32_t orc_mul_q15(int32_t s, int32_t g);    /* config.h:556-567      */
int32_t orc_f2i_sat(float x);

/* ---- parameter -> coefficient (host side, glibc libm) --------------------- */
void orc_eq_coeffs_f32(orc_eq_param *p, orc_biquad_f32 *bq, float fs);  /* dsp_pipeline.c:61-175 (float store) */
void orc_eq_coeffs_q28(orc_eq_param *p, orc_biquad_q28 *bq, float fs);  /* dsp_pipeline.c:61-175 (Q28 store)   */
void orc_xfeed_coeffs_f32(orc_xfeed_f32 *st, const orc_xfeed_cfg *cfg, float fs);  /* crossfeed.c:35-127 */
void orc_xfeed_coeffs_q28(orc_xfeed_q28 *st, const orc_xfeed_cfg *cfg, float fs);
void orc_lev_coeffs_compute(orc_lev_coeffs *out, const orc_lev_cfg *cfg, float fs); /* leveller.c:42-89 */
void orc_lev_reset_f32(orc_lev_state_f32 *st);                                     /* leveller.c:95-105 */
void orc_lev_reset_q28(orc_lev_state_q28 *st);
void orc_loud_table_f32(orc_loud_f32 table[ORC_LOUD_STEPS][2], float ref_spl, float intensity_pct, float fs); /* loudness.c:169-217 */
void orc_loud_table_q28(orc_loud_q28 table[ORC_LOUD_STEPS][2], float ref_spl, float intensity_pct, float fs);
int32_t orc_delay_samples(float delay_ms, float fs, int is_last, int32_t max_delay);  /* dsp_pipeline.c:216-239 */
int16_t orc_host_vol_mul(int16_t volume_8_8, uint8_t *vol_index_out);                /* usb_audio.c:410-440 */

/* ---- signal path: EQ cascades -------------------------------------------- */
/* dsp_pipeline.c:281-365 (block form; per-sample twin :256-279 gives the same values) */
void orc_f32s_eq_block(orc_biquad_f32 *bq, float *samples, uint32_t count, uint32_t nbands);
void orc_f32f_eq_block(orc_biquad_f32 *bq, float *samples, uint32_t count, uint32_t nbands);
/* dsp_process_rp2040.S:225-394 */
void orc_q28_eq_block(orc_biquad_q28 *bq, int32_t *samples, uint32_t count, uint32_t nbands);

/* many-channel drivers: bq[C][ORC_MAX_BANDS], samples[C][T]; processed in
 * packets of `packet` samples like the firmware (results do not depend on it) */
void orc_f32s_eq_many(orc_biquad_f32 *bq, float *samples, uint32_t C, uint32_t T, uint32_t nbands, uint32_t packet);
void orc_f32f_eq_many(orc_biquad_f32 *bq, float *samples, uint32_t C, uint32_t T, uint32_t nbands, uint32_t packet);
void orc_q28_eq_many(orc_biquad_q28 *bq, int32_t *samples, uint32_t C, uint32_t T, uint32_t nbands, uint32_t packet);

/* ---- signal path: stages -------------------------------------------------- */
void orc_f32s_xfeed(orc_xfeed_f32 *st, float *l, float *r, uint32_t count);   /* crossfeed.c:132-156 */
void orc_f32f_xfeed(orc_xfeed_f32 *st, float *l, float *r, uint32_t count);
void orc_q28_xfeed(orc_xfeed_q28 *st, int32_t *l, int32_t *r, uint32_t count); /* crossfeed.c:161-180 */
void orc_f32s_leveller(orc_lev_state_f32 *st, const orc_lev_coeffs *c, int lookahead, float *l, float *r, uint32_t count);  /* leveller.c:148-262 */
void orc_f32f_leveller(orc_lev_state_f32 *st, const orc_lev_coeffs *c, int lookahead, float *l, float *r, uint32_t count);
void orc_q28_leveller(orc_lev_state_q28 *st, const orc_lev_coeffs *c, int lookahead, int32_t *l, int32_t *r, uint32_t count); /* leveller.c:275-389 */
/* ---- S/PDIF subframe encoder (orc_spdif.c; pico_audio_spdif_multi) ---- */
void orc_spdif_lookup_init(uint32_t table[256]);                                                  /* audio_spdif.c:141-153 */
void orc_spdif_update_subframe(const uint32_t table[256], uint32_t *l, uint32_t *h, int32_t sample);  /* sample_encoding.h:27-50 */
void orc_spdif_encode(const uint32_t table[256], const int32_t *words, uint32_t frames, uint32_t pos0, const uint8_t cs[5], uint32_t *out);

/* usb_audio.c:459-498 update_preset_mute_envelope(): one packet of `sample_count` frames; returns the gain */
float orc_mute_envelope(uint8_t *preset_loading, uint32_t *preset_mute_counter, float *smooth_gain,
                        uint32_t sample_count, uint32_t sample_rate_hz);

void orc_pdm_reset(orc_pdm_state *st);
void orc_pdm_modulate(orc_pdm_state *st, int32_t sample_q28, uint32_t out[8]);  /* pdm_generator.c:351-397 */

/* ---- signal path: whole packet (usb_audio.c:500-1317, single-core branch) -- */
/* data: interleaved little-endian PCM (s16: 4 B/frame, s24: 6 B/frame).
 * spdif_out[pair][frame*2 + {0,1}] with pair stride `spdif_stride` words;
 * pdm_out[frame*8 .. +8] (only written when the sub output is enabled).
 * Returns number of frames processed. */
uint32_t orc_f32s_chain_packet(orc_chain_f32 *in, const uint8_t *data, uint32_t data_len, uint32_t bit_depth,
                               int32_t *spdif_out, uint32_t spdif_stride, uint32_t *pdm_out);
uint32_t orc_f32f_chain_packet(orc_chain_f32 *in, const uint8_t *data, uint32_t data_len, uint32_t bit_depth,
                               int32_t *spdif_out, uint32_t spdif_stride, uint32_t *pdm_out);
uint32_t orc_q28_chain_packet(orc_chain_q28 *in, const uint8_t *data, uint32_t data_len, uint32_t bit_depth,
                              int32_t *spdif_out, uint32_t spdif_stride, uint32_t *pdm_out);

#ifdef __cplusplus
}
#endif
#endif
