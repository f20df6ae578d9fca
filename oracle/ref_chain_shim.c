/*
 * ref_chain_shim.c — runs the reference's OWN process_audio_packet() (usb_audio.c:500-1317) on the host.
 *
 * TEST INFRASTRUCTURE (oracle/_ref only).  usb_audio.c is #included below from where it lies under
 * /root/reference, unmodified, once per platform build (-DPICO_RP2350=1 float / =0 Q28), and linked with
 * the reference's dsp_pipeline.c, crossfeed.c, leveller.c and loudness.c.  The pico-extras headers it
 * includes come from the reference tree; the SDK platform layer is oracle/stubs_fw/fw_stub.h.  This file
 * adds only
 *   (1) the neighbours process_audio_packet() touches: buffer pools whose buffers are the caller's capture
 *       memory, pdm_push_sample() recording the Q28 sub stream, a clock, the core-1 handshake globals;
 *   (2) a loader that copies one oracle instance record (orc_chain_f32 / orc_chain_q28, dspi_oracle.h)
 *       into the firmware's globals and the globals' state back afterwards — so that tests can run the same
 *       record through the restatement (orc_*_chain_packet) and through the reference and compare every
 *       byte of state and output;
 *   (3) for the RP2040 build, dsp_process_channel_block: the firmware's is Thumb assembly
 *       (dsp_process_rp2040.S); here it is the block loop around the reference's own compiled
 *       fast_mul_q28(), as in ref_shim.c;
 *   (4) for the RP2350 build, the two ARMv8-M inline-asm blocks of the 24-bit unpack (usb_audio.c:613-643,
 *       :657-674) cannot assemble on x86: a macro on __asm__ (defined AFTER every header is in) puts a C
 *       statement of what those instructions compute (sbfx/sxth/bfi/asr + vcvt.f32.s32) in front of them
 *       and leaves the asm itself in dead code.  16-bit input never reaches them.
 * Everything else that executes — unpack, preamp, loudness loops, master EQ call, leveller call, crossfeed
 * + input peaks, matrix, per-output EQ, gain, delay ring, peaks/clip, int24 conversion, PDM feed, the
 * preset-mute envelope — is the reference's compiled code.
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <xmmintrin.h>
#include <pmmintrin.h>
#include "dspi_oracle.h"

/* every header usb_audio.c includes, first: their guards keep them out of reach of the __asm__ macro below */
#include "pico/stdlib.h"
#include "pico/usb_device.h"
#include "pico/usb_device_private.h"
#include "pico/audio.h"
#include "pico/audio_spdif.h"
#include "pico/audio_i2s_multi.h"
#include "usb_audio.h"
#include "usb_descriptors.h"
#include "dsp_pipeline.h"
#include "dcp_inline.h"
#include "pdm_generator.h"
#include "flash_storage.h"
#include "loudness.h"
#include "crossfeed.h"
#include "leveller.h"
#include "bulk_params.h"
#include "pico/usb_stream_helper.h"
#include "usb_audio_ring.h"
#include "usb_feedback_controller.h"

#if PICO_RP2350
/* (4) what usb_audio.c:613-633 computes: three LE words -> L1 R1 L2 R2, sign-extended 24-bit, as float */
static inline void ref_unpack24(int32_t i0, int32_t i1, int32_t i2, float *l1, float *r1, float *l2, float *r2)
{
    int32_t t;
    t = (int32_t)((uint32_t)i0 << 8) >> 8;                                   /* sbfx temp, i0, #0, #24 */
    *l1 = (float)t;                                                          /* vcvt.f32.s32 */
    t = (int32_t)((uint32_t)(int32_t)(int16_t)i1 << 8);                      /* sxth; lsl #8 */
    t = (int32_t)(((uint32_t)t & ~0xFFu) | ((uint32_t)(i0 >> 24) & 0xFFu));  /* asr i0,#24; bfi temp,i0,#0,#8 */
    *r1 = (float)t;
    t = i1 >> 8;                                                             /* asr temp, i1, #8 */
    t = (int32_t)(((uint32_t)t & 0x00FFFFFFu) | ((uint32_t)i2 << 24));       /* bfi temp, i2, #24, #8 */
    t >>= 8;                                                                 /* asr temp, temp, #8 */
    *l2 = (float)t;
    t = i2 >> 8;                                                             /* asr temp, i2, #8 */
    *r2 = (float)t;
}
/* the odd-frame block (:651-677) declares only i0, i1, l1, r1: these file-scope names stand in there */
static float l2, r2;
static int32_t i2;
#define __asm__ ref_unpack24(i0, i1, i2, &l1, &r1, &l2, &r2); if (0) __asm__
#endif

#include "usb_audio.c"

#if PICO_RP2350
#undef __asm__
#endif

/* ---- (1) neighbours ------------------------------------------------------------------------------- */
volatile Core1Mode core1_mode = CORE1_MODE_PDM;           /* pdm_generator.c:41; PDM mode = single-core branch */
Core1EqWork core1_eq_work;                                /* pdm_generator.c:42 */
volatile bool pdm_enabled = false;                        /* pdm_generator.c:38 (only gates watermark stats) */
volatile bool preset_loading = false;                     /* flash_storage.c:255 */
volatile uint32_t preset_mute_counter = 0;                /* flash_storage.c:256 */
volatile uint32_t spdif_overruns, spdif_underruns, pdm_ring_overruns, pdm_ring_underruns, pdm_dma_overruns, pdm_dma_underruns;
pio_hw_t ref_pio_hw[3];

static uint32_t fake_us;
uint32_t time_us_32(void) { return fake_us += 7; }
uint64_t time_us_64(void) { return fake_us += 7; }
void ref_fw_event(void) { }
dma_hw_t *ref_dma_hw(void) { static dma_hw_t hw; return &hw; }

static int32_t *sub_capture;
static uint32_t sub_count;
void pdm_push_sample(int32_t sample, bool reset)          /* pdm_generator.c:186 — records instead of queueing */
{
    (void)reset;
    if (sub_capture) sub_capture[sub_count] = sample;
    sub_count++;
}

#define N_POOLS 4
static struct audio_buffer_pool pools[N_POOLS];
static struct audio_buffer pool_buf[N_POOLS];
static mem_buffer_t pool_mem[N_POOLS];
struct audio_buffer *take_audio_buffer(struct audio_buffer_pool *ac, bool block)
{
    (void)block;
    return &pool_buf[ac - pools];
}
void give_audio_buffer(struct audio_buffer_pool *ac, struct audio_buffer *buffer) { (void)ac; (void)buffer; }

/* ---- (3) RP2040: block cascade around the reference's compiled fast_mul_q28 -------------------------- */
#if !PICO_RP2350
void dsp_process_channel_block(Biquad *restrict bq, int32_t *restrict s, uint32_t count, uint8_t channel)
{
    const uint8_t nb = channel_band_counts[channel];                  /* dsp_process_rp2040.S:233-236 */
    for (uint8_t b = 0; b < nb; b++, bq++) {
        if (bq->bypass) continue;                                     /* :246-248 */
        int32_t s1 = bq->s1, s2 = bq->s2;
        for (uint32_t i = 0; i < count; i++) {
            int32_t x = s[i];
            int32_t y = (int32_t)((uint32_t)fast_mul_q28(bq->b0, x) + (uint32_t)s1);                          /* :272-290 */
            s1 = (int32_t)((uint32_t)fast_mul_q28(bq->b1, x) - (uint32_t)fast_mul_q28(bq->a1, y) + (uint32_t)s2);  /* :291-330 */
            s2 = (int32_t)((uint32_t)fast_mul_q28(bq->b2, x) - (uint32_t)fast_mul_q28(bq->a2, y));             /* :331-353 */
            s[i] = y;
        }
        bq->s1 = s1; bq->s2 = s2;
    }
}
int32_t dsp_process_channel(Biquad *restrict bq, int32_t x, uint8_t channel)
{
    dsp_process_channel_block(bq, &x, 1, channel);
    return x;
}
#endif

/* ---- (2) loader ------------------------------------------------------------------------------------- */
#if PICO_RP2350
typedef orc_chain_f32 chain_t;
#else
typedef orc_chain_q28 chain_t;
#endif
static LoudnessCoeffs loud_rows[LOUDNESS_BIQUAD_COUNT];

_Static_assert(sizeof(((chain_t *)0)->filters[0][0]) == sizeof(Biquad), "Biquad layout");
_Static_assert(sizeof(((chain_t *)0)->xfeed) == sizeof(CrossfeedState), "CrossfeedState layout");
_Static_assert(sizeof(((chain_t *)0)->levc) == sizeof(LevellerCoeffs), "LevellerCoeffs layout");
_Static_assert(sizeof(((chain_t *)0)->levs) == sizeof(LevellerState), "LevellerState layout");
_Static_assert(sizeof(((chain_t *)0)->loud[0]) == sizeof(LoudnessCoeffs), "LoudnessCoeffs layout");
_Static_assert(sizeof(orc_crosspoint) == sizeof(MatrixCrosspoint) && sizeof(orc_output) == sizeof(OutputChannel), "matrix layout");
_Static_assert(ORC_MAX_BANDS == MAX_BANDS && ORC_PKT_MAX == 192, "shape constants");

size_t ref_chain_sizeof(int which)
{
    switch (which) {
    case 0: return sizeof(chain_t);
    case 1: return NUM_OUTPUT_CHANNELS;
    case 2: return NUM_CHANNELS;
    case 3: return MAX_DELAY_SAMPLES;
    case 4: return NUM_SPDIF_INSTANCES;
    default: return 0;
    }
}

static void load(const chain_t *c, uint32_t fs, uint32_t bit_depth)
{
    audio_state.freq = fs;
    audio_state.vol_mul = c->host_vol_mul;
    audio_state.mute = c->host_mute;
    usb_input_bit_depth = (uint8_t)bit_depth;
    bypass_master_eq = c->bypass_master_eq;
#if PICO_RP2350
    master_volume_linear = c->master_volume_linear;
    global_preamp_linear[0] = c->preamp_linear[0]; global_preamp_linear[1] = c->preamp_linear[1];
    memcpy(loudness_state, c->loud_state, sizeof loudness_state);
#else
    master_volume_q15 = c->master_volume_q15;
    global_preamp_mul[0] = c->preamp_q28[0]; global_preamp_mul[1] = c->preamp_q28[1];
    for (int s = 0; s < 2; s++) for (int j = 0; j < LOUDNESS_BIQUAD_COUNT; j++) {
        loudness_biquads[s][j].s1 = c->loud_state[s][j].s1;
        loudness_biquads[s][j].s2 = c->loud_state[s][j].s2;
    }
#endif
    loudness_enabled = c->loudness_on;
    memcpy(loud_rows, c->loud, sizeof loud_rows);
    current_loudness_coeffs = loud_rows;
    crossfeed_bypassed = !c->crossfeed_on;
    memcpy(&crossfeed_state, &c->xfeed, sizeof crossfeed_state);
    leveller_bypassed = !c->leveller_on;
    memcpy(&leveller_coeffs, &c->levc, sizeof leveller_coeffs);
    leveller_config.enabled = c->leveller_on;
    leveller_config.lookahead = c->lev_lookahead;
    memcpy(&leveller_state, &c->levs, sizeof leveller_state);
    for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) {
        memcpy(&matrix_mixer.crosspoints[0][o], &c->xp[0][o], sizeof(MatrixCrosspoint));
        memcpy(&matrix_mixer.crosspoints[1][o], &c->xp[1][o], sizeof(MatrixCrosspoint));
        memcpy(&matrix_mixer.outputs[o], &c->out[o], sizeof(OutputChannel));
        channel_delay_samples[o] = c->delay_samples[o];
        memcpy(delay_lines[o], c->delay_lines[o], sizeof delay_lines[o]);
    }
    any_delay_active = c->any_delay_active;
    delay_write_idx = c->delay_widx;
    for (int ch = 0; ch < NUM_CHANNELS; ch++) {
        memcpy(filters[ch], c->filters[ch], sizeof filters[ch]);
        channel_bypassed[ch] = c->channel_bypassed[ch];
        global_status.peaks[ch] = c->peaks[ch];
    }
    global_status.clip_flags = c->clip_flags;
}

static void store(chain_t *c)
{
#if PICO_RP2350
    memcpy(c->loud_state, loudness_state, sizeof loudness_state);
#else
    for (int s = 0; s < 2; s++) for (int j = 0; j < LOUDNESS_BIQUAD_COUNT; j++) {
        c->loud_state[s][j].s1 = loudness_biquads[s][j].s1;
        c->loud_state[s][j].s2 = loudness_biquads[s][j].s2;
    }
#endif
    memcpy(&c->xfeed, &crossfeed_state, sizeof crossfeed_state);
    memcpy(&c->levs, &leveller_state, sizeof leveller_state);
    for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(c->delay_lines[o], delay_lines[o], sizeof delay_lines[o]);
    c->delay_widx = delay_write_idx;
    for (int ch = 0; ch < NUM_CHANNELS; ch++) {
        memcpy(c->filters[ch], filters[ch], sizeof filters[ch]);
        c->peaks[ch] = global_status.peaks[ch];
    }
    c->clip_flags = global_status.clip_flags;
}

/* preset-mute envelope state (usb_audio.c:457, flash_storage.c:255-256) */
void ref_chain_set_mute_env(int loading, uint32_t counter, float smooth_gain)
{
    preset_loading = loading != 0;
    preset_mute_counter = counter;
    preset_mute_smooth_gain = smooth_gain;
}
void ref_chain_get_mute_env(int *loading, uint32_t *counter, float *smooth_gain)
{
    *loading = preset_loading; *counter = preset_mute_counter; *smooth_gain = preset_mute_smooth_gain;
}

/* One USB packet through the reference.  Same contract as orc_*_chain_packet (dspi_oracle.h) except that
 * the sub output is returned as the Q28 samples handed to pdm_push_sample() (sub_q28[frame], *n_sub of them)
 * instead of modulated words (ref_pdm_shim.c runs the reference's modulator on them).  Preset-mute gain: with
 * the record's mute_env_on set, the record's envelope state is installed and the reference's
 * update_preset_mute_envelope() runs as in the firmware (state and resulting gain are written back); otherwise
 * the record's constant preset_mute_gain must be 0 or 1 and is installed as a settled envelope (the firmware
 * has no other way to produce a constant; returns 0xFFFFFFFF for anything else). */
uint32_t ref_chain_packet(chain_t *c, uint32_t fs, const uint8_t *data, uint32_t data_len, uint32_t bit_depth,
                          int32_t *spdif_out, uint32_t spdif_stride, int32_t *sub_q28, uint32_t *n_sub)
{
    if ((int)c->n_out != NUM_OUTPUT_CHANNELS || (int)c->max_delay != MAX_DELAY_SAMPLES) return 0xFFFFFFFEu;
    if (c->mute_env_on) {
        ref_chain_set_mute_env(c->preset_loading, c->preset_mute_counter, c->preset_mute_smooth_gain);
        fs = c->sample_rate_hz;
    } else {
        if (c->preset_mute_gain == 1.0f)      ref_chain_set_mute_env(0, 0, 1.0f);
        else if (c->preset_mute_gain == 0.0f) ref_chain_set_mute_env(1, 0x7FFFFFFFu, 0.0f);
        else return 0xFFFFFFFFu;
    }
    unsigned csr = _mm_getcsr();
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);                       /* firmware: FPSCR FZ (main.c:593-600) */
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    load(c, fs, bit_depth);
    static int32_t scratch[N_POOLS][192 * 2];
    producer_pool_1 = &pools[0];
    producer_pool_2 = &pools[1];
#if PICO_RP2350
    producer_pool_3 = &pools[2];
    producer_pool_4 = &pools[3];
#endif
    for (int p = 0; p < N_POOLS; p++) {
        pool_mem[p].bytes = (uint8_t *)scratch[p];
        pool_mem[p].size = sizeof scratch[p];
        pool_buf[p].buffer = &pool_mem[p];
        pool_buf[p].max_sample_count = 192;
    }
    sub_capture = sub_q28;
    sub_count = 0;
    process_audio_packet(data, (uint16_t)data_len);
    const uint32_t n = pool_buf[0].sample_count;
    for (int p = 0; p < NUM_SPDIF_INSTANCES; p++)
        memcpy(spdif_out + (size_t)p * spdif_stride, scratch[p], (size_t)n * 8);
    store(c);
    if (c->mute_env_on) {
        c->preset_loading = preset_loading;
        c->preset_mute_counter = preset_mute_counter;
        c->preset_mute_smooth_gain = preset_mute_smooth_gain;
        c->preset_mute_gain = preset_mute_smooth_gain;       /* the value update_preset_mute_envelope() returned */
    }
    if (n_sub) *n_sub = sub_count;
    _mm_setcsr(csr);
    return n;
}

/* the reference's audio_set_volume() (usb_audio.c:428-440): volume -> vol_mul and the loudness row index */
int16_t ref_host_vol_mul(int16_t volume_8_8, uint8_t *vol_index_out)
{
    loudness_enabled = true;
    loudness_active_table = loudness_tables[0];
    audio_set_volume(volume_8_8);
    if (vol_index_out) *vol_index_out = (uint8_t)((current_loudness_coeffs - &loudness_tables[0][0][0]) / LOUDNESS_BIQUAD_COUNT);
    return audio_state.vol_mul;
}

/* the reference's update_preamp() / update_master_volume() (usb_audio.c:244-269) */
void ref_preamp(float db, float *linear, int32_t *q28)
{
    update_preamp(0, db);
    *linear = global_preamp_linear[0];
    *q28 = global_preamp_mul[0];
}
void ref_master_volume(float db, float *linear, int32_t *q15)
{
    update_master_volume(db);
    *linear = master_volume_linear;
    *q15 = master_volume_q15;
}
