/*
 * dspi_b200.h — C ABI of the B200-native DSPi signal-chain engine.
 *
 * Drop-in boundary for the per-sample DSP hot path of WeebLabs/DSPi
 * (SURVEY.md §8b).  The reference has no FFI layer: the path is reached through
 * plain C functions over global arrays.  Every entry point below names the
 * reference function (file:line under /root/reference/firmware/DSPi) whose role
 * it takes for MANY independent channels / device instances at once.
 *
 * Conventions
 *   - plain pointers and sizes only; records are byte-for-byte the reference's
 *     (sizes asserted below; checked against the compiled reference in tests);
 *   - every function returns 0 on success or a negative DSPI_E* code, never
 *     aborts; dspi_last_error() gives a per-thread message;
 *   - an engine belongs to one CUDA device and one stream; calls on one engine
 *     must be serialised by the caller (the firmware's single processing
 *     thread, main.c:743), different engines are independent;
 *   - there is NO CPU fallback: if no sm_100 device is present, create fails
 *     with DSPI_ENODEV.
 */
#ifndef DSPI_B200_H
#define DSPI_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSPI_MAX_BANDS        12   /* config.h:329 MAX_BANDS — row stride of filters[][]          */
#define DSPI_NUM_BANDS        10   /* dsp_pipeline.c:36-44 channel_band_counts                     */
#define DSPI_LA_SAMPLES      480   /* leveller.h:36                                                */
#define DSPI_PACKET_MAX      192   /* usb_audio.c:273,588                                          */

enum {
    DSPI_OK       = 0,
    DSPI_EINVAL   = -22,   /* bad argument                                   */
    DSPI_ENOMEM   = -12,   /* host or device allocation failed               */
    DSPI_ENODEV   = -19,   /* no sm_100 CUDA device / CUDA runtime unusable  */
    DSPI_ECUDA    = -5,    /* a CUDA call failed (message in dspi_last_error) */
    DSPI_ERANGE   = -34    /* index / size outside the engine's shape        */
};

/* arithmetic of an engine (SURVEY.md §8c "Build flags for the oracle") */
enum {
    DSPI_ARITH_F32_FUSED  = 0,  /* RP2350 float path as arm-none-eabi-gcc contracts it (VFMA) */
    DSPI_ARITH_F32_STRICT = 1,  /* RP2350 float path, every operation rounded separately       */
    DSPI_ARITH_Q28        = 2   /* RP2040 Q28 fixed point (dsp_process_rp2040.S), bit-exact    */
};

/* filter types, config.h:440-443 */
enum { DSPI_FILTER_FLAT = 0, DSPI_FILTER_PEAKING = 1, DSPI_FILTER_LOWSHELF = 2,
       DSPI_FILTER_HIGHSHELF = 3, DSPI_FILTER_LOWPASS = 4, DSPI_FILTER_HIGHPASS = 5 };

/* ---- records shared with the reference (same bytes) ----------------------- */

/* Biquad, RP2350 build — config.h:418-431 (68 bytes) */
typedef struct {
    float b0, b1, b2, a1, a2;
    float s1, s2;
    float sva1, sva2, sva3;
    float svm0, svm1, svm2;
    float svic1eq, svic2eq;
    uint32_t svf_type;
    uint8_t use_svf;
    uint8_t bypass;
} dspi_biquad_f32;

/* Biquad, RP2040 build — config.h:433-437, dsp_process_rp2040.S:6-14 (32 bytes) */
typedef struct {
    int32_t b0, b1, b2, a1, a2;
    int32_t s1, s2;
    uint8_t bypass;
} dspi_biquad_q28;

/* EqParamPacket — config.h:445-453 (packed, 16 bytes) */
typedef struct __attribute__((packed)) {
    uint8_t channel, band, type, reserved;
    float freq, Q, gain_db;
} dspi_eq_param;

#ifdef __cplusplus
static_assert(sizeof(dspi_biquad_f32) == 68 && sizeof(dspi_biquad_q28) == 32 && sizeof(dspi_eq_param) == 16, "reference layouts");
#else
_Static_assert(sizeof(dspi_biquad_f32) == 68 && sizeof(dspi_biquad_q28) == 32 && sizeof(dspi_eq_param) == 16, "reference layouts");
#endif

/* ---- library ---------------------------------------------------------------- */
const char *dspi_last_error(void);
/* number of visible CUDA devices that can run the engine (compute capability 10.x) */
int dspi_device_count(void);

/* ---- host-side parameter API (no GPU needed) -------------------------------- */
/* dsp_compute_coefficients(), dsp_pipeline.c:61-175.  `p` is clamped in place
 * exactly like the reference does (:78-81).  The float store keeps SVF/biquad
 * path selection and the state reset on a path flip (:87-92). */
void dspi_compute_coefficients_f32(dspi_eq_param *p, dspi_biquad_f32 *bq, float sample_rate);
void dspi_compute_coefficients_q28(dspi_eq_param *p, dspi_biquad_q28 *bq, float sample_rate);

/* ---- EQ engine: many independent cascades ---------------------------------- */
/* One row of filters[][] per channel: dsp_process_channel_block()
 * (float dsp_pipeline.c:281-365, Q28 dsp_process_rp2040.S:225-394) applied to
 * n_channels channels in one launch. */
typedef struct dspi_eq dspi_eq;

typedef struct {
    uint32_t arith;          /* DSPI_ARITH_*                                          */
    uint32_t n_channels;     /* independent EQ channels (rows of filters[][])         */
    uint32_t n_bands;        /* bands processed per channel, 1..DSPI_MAX_BANDS (10)   */
    int32_t  device;         /* CUDA device ordinal                                   */
    uint32_t flags;          /* 0                                                     */
} dspi_eq_desc;

int dspi_eq_create(dspi_eq **out, const dspi_eq_desc *desc);
int dspi_eq_destroy(dspi_eq *e);

/* Coefficients + state in the reference's own layout:
 * biquads[n][DSPI_MAX_BANDS] of dspi_biquad_f32 (float engines) or
 * dspi_biquad_q28 (Q28 engines) — i.e. n rows of the firmware's filters[][].
 * upload == writing filters[][] between packets (main.c:843-856 semantics:
 * takes effect for the next process call); download returns coefficients and
 * the CURRENT filter state, so a run can be checkpointed bit-exactly. */
int dspi_eq_upload_biquads(dspi_eq *e, uint32_t ch0, uint32_t n, const void *biquads);
int dspi_eq_download_biquads(dspi_eq *e, uint32_t ch0, uint32_t n, void *biquads);
/* REQ_SET_EQ_PARAM path (usb_audio.c:1641-1649 + main.c:826-857): recompute one
 * band of one channel from its recipe and upload it.  p->channel is ignored
 * (8 bits on the wire); `channel` selects the row, p->band the band. */
int dspi_eq_set_param(dspi_eq *e, uint32_t channel, dspi_eq_param *p, float sample_rate);
/* Mass reconfiguration (device part): dsp_compute_coefficients() for all 12 bands of channels
 * [ch0, ch0 + n) ON THE GPU, straight into the engine's stores - recipes[n][DSPI_MAX_BANDS] (host memory) are
 * clamped in place like the reference does (dsp_pipeline.c:78-81); filter state is kept unless a band's
 * topology flips (:87-92).  Arithmetic is the reference's float arithmetic operation by operation; its libm
 * calls (powf, tanf, sinf, cosf) are evaluated in double and rounded once ("libm policy"), which
 * can differ from a host libm's float functions in a last bit - use dspi_compute_coefficients_* +
 * dspi_eq_upload_biquads when coefficients must be those of a particular host libm. */
int dspi_eq_set_params_device(dspi_eq *e, uint32_t ch0, uint32_t n, dspi_eq_param *recipes, float sample_rate);

/* Process T samples of every channel, in place.
 *   *_device: samples is a DEVICE pointer, channel-major [n_channels][ld]
 *             (row stride `ld` elements, float32 or int32); asynchronous on the
 *             engine's stream.
 *   *_host:   samples is a HOST pointer [n_channels][T]; copies in, processes,
 *             copies out (chunked, overlapped) and returns when the data is back.
 *             Pinned memory (dspi_host_alloc) gives full PCIe bandwidth.
 * Equivalent reference loop: for each channel, dsp_process_channel_block(
 * filters[ch], samples[ch], T, ch) — packet size does not change the values. */
int dspi_eq_process_device(dspi_eq *e, void *d_samples, uint32_t T, uint32_t ld);
/* The same for channels [ch0, ch0 + n) only: d_rows points at the row of channel ch0 ([n][ld], device memory); ch0 must be a
 * multiple of 64.  For callers that stream a large block through in pieces (dspi_b200/sharding.py pipelines NCCL transfers
 * against it).  Asynchronous on the engine's stream. */
int dspi_eq_process_device_range(dspi_eq *e, void *d_rows, uint32_t T, uint32_t ld, uint32_t ch0, uint32_t n);
/* Host block: staged through the device in 48 MiB channel chunks (DSPI_HOST_CHUNK_MB in the environment overrides). */
int dspi_eq_process_host(dspi_eq *e, void *h_samples, uint32_t T);
int dspi_eq_sync(dspi_eq *e);
/* cudaStream_t of the engine (so callers can order their own work / events) */
void *dspi_eq_stream(dspi_eq *e);
/* number of kernel launches issued by this engine so far */
uint64_t dspi_eq_launch_count(dspi_eq *e);
/* Which K1/K2 kernel the next dspi_eq_process_* call will run, as text ("jit sig=0x...", "aot
 * straight-line biquad", "aot generic ...").  Float engines whose channels share one band-topology
 * vector get a kernel compiled for that vector at run time (NVRTC); this call triggers that
 * compilation if it is pending.  Results never depend on the choice.  No reference counterpart. */
int dspi_eq_kernel_info(dspi_eq *e, char *buf, size_t cap);

/* ---- EQ engine over several GPUs of one box, one process (SURVEY 8b `devices[], n_devices`, 8e) ------ */
/* Channels shard into contiguous ranges, one per listed device (dspi_eqx_shard_range: 64-channel boundaries); all
 * coefficients and filter state of a range stay on its owner and no data is exchanged between devices - the firmware's
 * own two-core split is the same idea (disjoint output ranges, config.h:350-357).  Results are bit-identical to a
 * single engine over all channels.  upload / download address channels of the whole group. */
#define DSPI_MAX_DEVICES 8
typedef struct dspi_eqx dspi_eqx;
typedef struct {
    uint32_t arith;                     /* DSPI_ARITH_*                                     */
    uint32_t n_channels;                /* over the whole group                             */
    uint32_t n_bands;
    uint32_t n_devices;                 /* 1..DSPI_MAX_DEVICES                              */
    int32_t  devices[DSPI_MAX_DEVICES]; /* CUDA ordinals; devices[0] is the root            */
    uint32_t flags;                     /* 0                                                */
} dspi_eqx_desc;
int dspi_eqx_create(dspi_eqx **out, const dspi_eqx_desc *desc);
int dspi_eqx_destroy(dspi_eqx *x);
int dspi_eqx_shard_range(uint32_t n_channels, uint32_t n_devices, uint32_t k, uint32_t *lo, uint32_t *hi);
int dspi_eqx_upload_biquads(dspi_eqx *x, uint32_t ch0, uint32_t n, const void *biquads);
int dspi_eqx_download_biquads(dspi_eqx *x, uint32_t ch0, uint32_t n, void *biquads);
/* samples [n_channels][T] in HOST memory (pinned: dspi_host_alloc): every device runs its staged PCIe pipeline on its
 * rows concurrently; returns when all rows are back. */
int dspi_eqx_process_host(dspi_eqx *x, void *h_samples, uint32_t T);
/* samples [n_channels][T] resident on devices[0]: the root processes its rows in place, the other devices pull theirs
 * over NVLink (peer access), process and push them back, chunked so that transfers in both directions overlap the
 * kernels.  Returns when the block on the root is complete. */
int dspi_eqx_process_root(dspi_eqx *x, void *d_samples_on_root, uint32_t T, uint32_t ld);
uint64_t dspi_eqx_launch_count(dspi_eqx *x);

/* ---- one rank per GPU: frames that originate on one rank, over NCCL ------------------------- */
/* Multi-PROCESS form of dspi_eqx_process_root: every rank owns an EQ engine over its shard
 * [lo_r, hi_r) = dspi_eqx_shard_range(total, world, r); the block [total][T] lives on the root rank.  dspi_sg_process
 * scatters it in row chunks with grouped ncclSend / ncclRecv: step j carries chunk j out and chunk j-L back in the same
 * NCCL group (both NVLink directions busy) while the engines work on the chunks in between, each chunk's kernel on a
 * stream of its own (a cascade kernel takes as long as its rows are long however few rows it gets, but a chunk fills only
 * a few SMs); the results end up in the root's block.  n_chunks = 0 lets the library choose chunk count and lag from the
 * transfer / kernel time ratio.  libnccl.so.2 is dlopen'ed at first use.  The 128-byte unique id comes from
 * dspi_nccl_unique_id() on one rank and is
 * handed to the others by the caller (dspi_b200/sharding.py broadcasts it with torch.distributed). */
typedef struct dspi_sg dspi_sg;
int dspi_nccl_unique_id(void *id128);
int dspi_sg_create(dspi_sg **out, dspi_eq *engine, int device, const void *id128, int rank, int world, int root);
int dspi_sg_destroy(dspi_sg *g);
int dspi_sg_process(dspi_sg *g, void *d_full_on_root, uint32_t total_channels, uint32_t T, uint32_t n_chunks);

/* ---- full signal chain: many independent DSPi device instances ------------------------------ */
/* One instance = process_audio_packet() of one RP2350-shape device (usb_audio.c:500-1317, float
 * pipeline :560-967, single-core branch :874-960): 2 inputs -> preamp -> loudness -> master EQ ->
 * leveller -> crossfeed -> 2x9 matrix -> per-output EQ, gain, delay -> 4 S/PDIF stereo pairs
 * (24-bit words) + the PDM sub through the 2nd-order delta-sigma modulator (pdm_generator.c:351-397). */
#define DSPI_CHAIN_OUTPUTS      9    /* config.h:321 NUM_OUTPUT_CHANNELS                            */
#define DSPI_CHAIN_EQ_CHANNELS 11    /* config.h:322 NUM_CHANNELS: master L, R, Out1..9             */
#define DSPI_CHAIN_MAX_DELAY 4096    /* config.h:84  MAX_DELAY_SAMPLES                              */

/* LoudnessCoeffs (RP2350), loudness.h:11-17 (28 bytes) */
typedef struct { float sva1, sva2, sva3, svm0, svm1, svm2; uint8_t bypass; } dspi_loudness_coeffs_f32;
/* CrossfeedState (RP2350), crossfeed.h:46-51 (28 bytes) */
typedef struct { float lp_a0, lp_b1, lp_state_L, lp_state_R, ap_a, ap_state_L, ap_state_R; } dspi_crossfeed_state_f32;
/* LevellerCoeffs, leveller.h:81-99 (36 bytes) */
typedef struct {
    float alpha_rms, alpha_attack, alpha_release;
    float threshold_db, ratio, knee_width_db, makeup_db, gate_threshold_db, max_gain_db;
} dspi_leveller_coeffs;
/* MatrixCrosspoint / OutputChannel / MatrixMixer, config.h:383-406 (12 / 20 / 396 bytes) */
typedef struct __attribute__((packed)) { uint8_t enabled, phase_invert, reserved[2]; float gain_db, gain_linear; } dspi_matrix_crosspoint;
typedef struct __attribute__((packed)) { uint8_t enabled, mute, reserved[2]; float gain_db, gain_linear, delay_ms; int32_t delay_samples; } dspi_output_channel;
typedef struct {
    dspi_matrix_crosspoint crosspoints[2][DSPI_CHAIN_OUTPUTS];
    dspi_output_channel outputs[DSPI_CHAIN_OUTPUTS];
} dspi_matrix_mixer_f32;
/* SystemStatusPacket, config.h:455-460 (26 bytes): peaks are Q15, clip_flags sticky */
typedef struct { uint16_t peaks[DSPI_CHAIN_EQ_CHANNELS]; uint8_t cpu0_load, cpu1_load; uint16_t clip_flags; } dspi_status;

/* Everything process_audio_packet() reads besides filters[][]: the globals of usb_audio.c:148-211
 * for one instance, with the coefficient records in the reference's own layouts. */
typedef struct {
    uint8_t bypass_master_eq;        /* usb_audio.c:48                                             */
    uint8_t loudness_enabled;        /* loud_on && current_loudness_coeffs != NULL, :579-580        */
    uint8_t crossfeed_enabled;       /* !crossfeed_bypassed                                         */
    uint8_t leveller_enabled;        /* !leveller_bypassed                                          */
    uint8_t host_mute;               /* audio_state.mute                                            */
    uint8_t leveller_lookahead;      /* leveller_config.lookahead                                   */
    uint8_t reserved0[2];
    int16_t host_vol_mul;            /* audio_state.vol_mul - an int16: 0 dB gives -32768 (quirk)   */
    int16_t reserved1;
    float preset_mute_gain;          /* update_preset_mute_envelope(), 1.0 when no preset loads     */
    float master_volume_linear;      /* usb_audio.c:161                                             */
    float preamp_linear[2];          /* global_preamp_linear[]                                      */
    dspi_loudness_coeffs_f32 loudness[2];   /* the selected row of loudness_active_table            */
    dspi_crossfeed_state_f32 crossfeed;     /* coefficients AND state (crossfeed_compute_coefficients clears state) */
    dspi_leveller_coeffs leveller;
    dspi_matrix_mixer_f32 matrix;    /* outputs[o].delay_samples is channel_delay_samples[o]        */
} dspi_chain_params_f32;

#ifdef __cplusplus
static_assert(sizeof(dspi_loudness_coeffs_f32) == 28 && sizeof(dspi_crossfeed_state_f32) == 28 && sizeof(dspi_leveller_coeffs) == 36 &&
              sizeof(dspi_matrix_mixer_f32) == 396 && sizeof(dspi_status) == 26, "reference layouts");
#else
_Static_assert(sizeof(dspi_loudness_coeffs_f32) == 28 && sizeof(dspi_crossfeed_state_f32) == 28 && sizeof(dspi_leveller_coeffs) == 36 &&
               sizeof(dspi_matrix_mixer_f32) == 396 && sizeof(dspi_status) == 26, "reference layouts");
#endif

/* CrossfeedConfig, crossfeed.h:26-32 (12 bytes) and LevellerConfig, leveller.h:59-66 (24 bytes) */
typedef struct { uint8_t enabled, itd_enabled, preset; float custom_fc, custom_feed_db; } dspi_crossfeed_config;
typedef struct { uint8_t enabled; float amount; uint8_t speed; float max_gain_db; uint8_t lookahead; float gate_threshold_db; } dspi_leveller_config;
/* host-side parameter functions of the chain (no GPU needed; host libm, like the firmware's main loop):
 * crossfeed_compute_coefficients() crossfeed.c:35-127 (clears the filter state),
 * leveller_compute_coefficients() leveller.c:42-89, loudness_recompute_table() loudness.c:169-217
 * (table[volume step 0..60][low shelf, high shelf]) and audio_set_volume() usb_audio.c:428-440
 * (returns audio_state.vol_mul and the loudness table row). */
void dspi_crossfeed_compute_coefficients_f32(dspi_crossfeed_state_f32 *st, const dspi_crossfeed_config *cfg, float sample_rate);
void dspi_leveller_compute_coefficients(dspi_leveller_coeffs *out, const dspi_leveller_config *cfg, float sample_rate);
void dspi_loudness_compute_table_f32(dspi_loudness_coeffs_f32 table[61][2], float ref_spl, float intensity_pct, float sample_rate);
int16_t dspi_host_volume(int16_t volume_8_8, uint8_t *table_index);
/* update_preamp() usb_audio.c:244-250 and update_master_volume() :255-269: dB -> the gains the packet loop reads
 * (float for the RP2350 shape, Q28 / Q15 for the RP2040 shape).  Return -1 for NaN / Inf like the firmware's guard. */
int dspi_preamp(float db, float *linear_out, int32_t *q28_out);
int dspi_master_volume(float db, float *linear_out, int32_t *q15_out);
/* The preset-mute envelope, update_preset_mute_envelope() usb_audio.c:456-498: state of one instance and its
 * per-packet step (host side; dspi_chain(q)_set_preset_mute runs the same recurrence on the device). */
typedef struct {
    uint8_t  loading;                /* preset_loading, flash_storage.c:255                          */
    uint8_t  reserved[3];
    uint32_t counter;                /* preset_mute_counter, flash_storage.c:256                     */
    float    smooth_gain;            /* preset_mute_smooth_gain, usb_audio.c:457 (1.0 = full level)  */
} dspi_preset_mute;
void  dspi_preset_mute_arm(dspi_preset_mute *m, uint32_t sample_rate_hz);    /* flash_storage.c:272-276, 347-348 */
float dspi_preset_mute_step(dspi_preset_mute *m, uint32_t sample_count, uint32_t sample_rate_hz);

typedef struct dspi_chain dspi_chain;
typedef struct {
    uint32_t arith;          /* DSPI_ARITH_F32_FUSED or DSPI_ARITH_F32_STRICT                       */
    uint32_t n_instances;
    uint32_t n_bands;        /* channel_band_counts[] value (10)                                    */
    int32_t  device;
    uint32_t max_frames;     /* largest n_packets * frames_per_packet of one process call           */
} dspi_chain_desc;

int dspi_chain_create(dspi_chain **out, const dspi_chain_desc *desc);
int dspi_chain_destroy(dspi_chain *c);
/* bulk_params_apply()-style update between packets (bulk_params.c:178-377 + main.c:1126-1162):
 * params[n] for instances [inst0, inst0+n).  Filter, delay-line, leveller and PDM state are kept; so is the crossfeed
 * filter state unless the record's crossfeed COEFFICIENTS differ from the ones in force - then the record's state rows
 * are taken, i.e. the zeros crossfeed_compute_coefficients() leaves (crossfeed.c:35-127 is the only place the firmware
 * resets that state; audio_set_volume and the mute / matrix handlers never do). */
int dspi_chain_set_params(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_chain_params_f32 *params);
/* filters[NUM_CHANNELS][MAX_BANDS] of n instances: biquads[n][11][12] (master L, R, Out1..9) */
int dspi_chain_upload_biquads(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_biquad_f32 *biquads);
int dspi_chain_download_biquads(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_biquad_f32 *biquads);
/* dsp_recalculate_all_filters() for n instances on the GPU: recipes[n][11][DSPI_MAX_BANDS] = filter_recipes[][] of each
 * instance (host memory, clamped in place); see dspi_eq_set_params_device for the arithmetic and the libm policy */
int dspi_chain_set_eq_params_device(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_eq_param *recipes, float sample_rate);
/* pipeline reset: clears leveller, loudness, delay-line and PDM state (leveller_reset_state(),
 * pdm_processing_loop() restart path); filter state is part of the biquads */
int dspi_chain_reset_state(dspi_chain *c);
/* The preset-mute envelope inside the engine (update_preset_mute_envelope(), usb_audio.c:466-498, called once per packet
 * at :532): states[n] puts instances [inst0, inst0+n) into envelope mode - from then on every packet of every process
 * call advances the instance's envelope and uses its gain where process_audio_packet() uses preset_mute_gain (:570), so a
 * fade runs across the packets of one call and across calls.  Arm a mute with dspi_preset_mute_arm() as the firmware's
 * flash operations do.  states == NULL leaves envelope mode: the constant preset_mute_gain of dspi_chain_set_params
 * applies again.  _get_ returns the current state (it is part of the state blob too). */
/* Mass reconfiguration of the dynamics stages ON THE GPU: per instance what the firmware's main loop does
 * when crossfeed_update_pending / leveller_update_pending / loudness_recompute_pending are set (main.c:868-895) -
 * crossfeed_compute_coefficients() (crossfeed.c:35-127: new coefficients, filter state cleared), leveller_compute_coefficients()
 * (leveller.c:42-89), loudness_recompute_table() (loudness.c:169-217) for the row audio_set_volume() selects - followed by
 * audio_set_volume() (usb_audio.c:428-440): the host volume becomes vol_mul (the int16 quirk included) and the output gains
 * follow.  cfgs[n] is host memory.  Arithmetic and libm policy as dspi_eq_set_params_device.  The bypass flags, preamp,
 * master volume, matrix and delays stay as dspi_chain_set_params left them. */
typedef struct {
    dspi_crossfeed_config crossfeed;     /* crossfeed_config, usb_audio.c:187-193                       */
    dspi_leveller_config  leveller;      /* leveller_config, usb_audio.c:199-206                        */
    float   loudness_ref_spl;            /* loudness_ref_spl, usb_audio.c:175                           */
    float   loudness_intensity_pct;      /* loudness_intensity_pct, usb_audio.c:176                     */
    uint8_t loudness_enabled;            /* loudness_enabled, usb_audio.c:174                           */
    uint8_t host_mute;                   /* audio_state.mute                                            */
    int16_t volume_8_8;                  /* audio_state.volume: UAC1 volume in 1/256 dB, 0 = full scale */
} dspi_dynamics_config;
int dspi_chain_set_dynamics_device(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_dynamics_config *cfgs, float sample_rate);
int dspi_chain_set_preset_mute(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_preset_mute *states, uint32_t sample_rate_hz);
int dspi_chain_get_preset_mute(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_preset_mute *states);
/* Checkpoint / resume (the dspi_state_export/import of SURVEY 8 b): everything a later process call depends on besides
 * dspi_chain_set_params' records - filter coefficients and state, loudness / crossfeed / leveller state, look-ahead and
 * delay rings, write index, modulator state, meters.  The blob is private to this library (header + raw arrays) and only
 * loads into an engine of the same shape. */
size_t dspi_chain_state_size(dspi_chain *c);
int dspi_chain_state_export(dspi_chain *c, void *blob, size_t cap);
int dspi_chain_state_import(dspi_chain *c, const void *blob, size_t len);
/* n_packets USB packets of frames_per_packet (<= 192) frames for every instance.
 *   pcm:       [n_instances][n_packets * frames_per_packet] interleaved L,R little-endian frames,
 *              bit_depth 16 (4 bytes / frame) or 24 (packed, 6 bytes / frame)      (HOST memory)
 *   spdif_out: [n_instances][4][n_frames][2] int32 - the four pico_audio producer buffers
 *   pdm_out:   [n_instances][n_frames][8] uint32 - 256 PDM bits per frame, MSB first (written only
 *              for instances whose sub output is enabled)
 *   status:    [n_instances], peaks of the LAST packet, clip flags OR-ed in (sticky)
 * Any of the three outputs may be NULL. */
int dspi_chain_process_host(dspi_chain *c, const void *pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t frames_per_packet,
                            int32_t *spdif_out, uint32_t *pdm_out, dspi_status *status);
/* same with DEVICE pointers; asynchronous on the engine stream */
int dspi_chain_process_device(dspi_chain *c, const void *d_pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t frames_per_packet,
                              int32_t *d_spdif_out, uint32_t *d_pdm_out, dspi_status *d_status);
int dspi_chain_sync(dspi_chain *c);
void *dspi_chain_stream(dspi_chain *c);
uint64_t dspi_chain_launch_count(dspi_chain *c);
/* How the engine split the GPU for this chain: the delta-sigma modulator (one serial chain per instance, latency-bound)
 * runs alone on pdm_sms SMs, every other stage on rest_sms (CUDA green contexts; 0 / 0 when the driver offers none or
 * DSPI_PDM_SMS=0 is set - results never depend on it).  No reference counterpart: the firmware gives the modulator
 * core 1 (pdm_generator.c:691-721). */
int dspi_chain_sm_partition(dspi_chain *c, uint32_t *pdm_sms, uint32_t *rest_sms);
/* dsp_update_delay_samples() for one output, dsp_pipeline.c:216-239 (is_last adds SUB_ALIGN_SAMPLES) */
int32_t dspi_delay_samples(float delay_ms, float sample_rate, int is_last);

/* ---- full signal chain, RP2040 arithmetic (Q28 fixed point, 2 in -> 5 out) ---------------------- */
/* process_audio_packet(), usb_audio.c:968-1283 (single-core branch :1191-1276): every stage in
 * 32-bit wrapping Q28/Q15 arithmetic (fast_mul_q28 dsp_pipeline.c:47-58, fast_mul_q15
 * config.h:556-567), crossfeed.c:161-180, leveller.c:275-389; bit-exact. */
#define DSPI_CHAINQ_OUTPUTS      5    /* config.h:326                                               */
#define DSPI_CHAINQ_EQ_CHANNELS  7    /* config.h:327                                               */
#define DSPI_CHAINQ_MAX_DELAY 2048    /* config.h:86                                                */

/* LoudnessCoeffs (RP2040), loudness.h:21 (24 bytes); CrossfeedState (RP2040), crossfeed.h:53-58 */
typedef struct { int32_t b0, b1, b2, a1, a2; uint8_t bypass; } dspi_loudness_coeffs_q28;
typedef struct { int32_t lp_a0, lp_b1, lp_state_L, lp_state_R, ap_a, ap_state_L, ap_state_R; } dspi_crossfeed_state_q28;
/* MatrixMixer (RP2040), config.h:403-406 (220 bytes); SystemStatusPacket (RP2040, 18 bytes) */
typedef struct {
    dspi_matrix_crosspoint crosspoints[2][DSPI_CHAINQ_OUTPUTS];
    dspi_output_channel outputs[DSPI_CHAINQ_OUTPUTS];
} dspi_matrix_mixer_q28;
typedef struct { uint16_t peaks[DSPI_CHAINQ_EQ_CHANNELS]; uint8_t cpu0_load, cpu1_load; uint16_t clip_flags; } dspi_status_q28;

typedef struct {
    uint8_t bypass_master_eq, loudness_enabled, crossfeed_enabled, leveller_enabled;
    uint8_t host_mute, leveller_lookahead, reserved0[2];
    int16_t host_vol_mul;            /* audio_state.vol_mul (int16: 0 dB gives -32768)              */
    int16_t reserved1;
    float preset_mute_gain;          /* quantised to Q15 per packet, usb_audio.c:976-978            */
    int32_t master_volume_q15;       /* usb_audio.c:162                                             */
    int32_t preamp_q28[2];           /* global_preamp_mul[]                                         */
    dspi_loudness_coeffs_q28 loudness[2];
    dspi_crossfeed_state_q28 crossfeed;
    dspi_leveller_coeffs leveller;
    dspi_matrix_mixer_q28 matrix;
} dspi_chain_params_q28;

#ifdef __cplusplus
static_assert(sizeof(dspi_loudness_coeffs_q28) == 24 && sizeof(dspi_crossfeed_state_q28) == 28 && sizeof(dspi_matrix_mixer_q28) == 220 &&
              sizeof(dspi_status_q28) == 18, "reference layouts");
#else
_Static_assert(sizeof(dspi_loudness_coeffs_q28) == 24 && sizeof(dspi_crossfeed_state_q28) == 28 && sizeof(dspi_matrix_mixer_q28) == 220 &&
               sizeof(dspi_status_q28) == 18, "reference layouts");
#endif

typedef struct dspi_chainq dspi_chainq;
int dspi_chainq_create(dspi_chainq **out, const dspi_chain_desc *desc);      /* desc->arith must be DSPI_ARITH_Q28 */
int dspi_chainq_destroy(dspi_chainq *c);
int dspi_chainq_set_params(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_chain_params_q28 *params);
/* filters[7][12] per instance: biquads[n][7][12] (master L, R, Out1..4, sub) */
int dspi_chainq_upload_biquads(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_biquad_q28 *biquads);
int dspi_chainq_download_biquads(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_biquad_q28 *biquads);
int dspi_chainq_set_eq_params_device(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_eq_param *recipes, float sample_rate);   /* recipes[n][7][12] */
int dspi_chainq_reset_state(dspi_chainq *c);
int dspi_chainq_set_dynamics_device(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_dynamics_config *cfgs, float sample_rate);
int dspi_chainq_set_preset_mute(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_preset_mute *states, uint32_t sample_rate_hz);   /* Q15 use of the gain: usb_audio.c:976-980 */
int dspi_chainq_get_preset_mute(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_preset_mute *states);
size_t dspi_chainq_state_size(dspi_chainq *c);
int dspi_chainq_state_export(dspi_chainq *c, void *blob, size_t cap);
int dspi_chainq_state_import(dspi_chainq *c, const void *blob, size_t len);
/* pcm as for dspi_chain_process_host; spdif_out [n_instances][2][n_frames][2]; pdm_out [n_instances][n_frames][8] */
int dspi_chainq_process_host(dspi_chainq *c, const void *pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t frames_per_packet,
                             int32_t *spdif_out, uint32_t *pdm_out, dspi_status_q28 *status);
int dspi_chainq_process_device(dspi_chainq *c, const void *d_pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t frames_per_packet,
                               int32_t *d_spdif_out, uint32_t *d_pdm_out, dspi_status_q28 *d_status);
int dspi_chainq_sync(dspi_chainq *c);
void *dspi_chainq_stream(dspi_chainq *c);
uint64_t dspi_chainq_launch_count(dspi_chainq *c);
int dspi_chainq_sm_partition(dspi_chainq *c, uint32_t *pdm_sms, uint32_t *rest_sms);
/* Q28 stores of the crossfeed and loudness parameter functions (crossfeed.c:116-119, loudness.c:131-162) */
void dspi_crossfeed_compute_coefficients_q28(dspi_crossfeed_state_q28 *st, const dspi_crossfeed_config *cfg, float sample_rate);
void dspi_loudness_compute_table_q28(dspi_loudness_coeffs_q28 table[61][2], float ref_spl, float intensity_pct, float sample_rate);

/* ---- bulk parameter ingest (host side; SURVEY.md 8 f-1 / the dspi_set_bulk_params of 8 b) ------- */
/* WireBulkParams, bulk_params.h:40-205: the 2896-byte little-endian packet the DSPi Console sends with
 * REQ_SET_ALL_PARAMS and reads with REQ_GET_ALL_PARAMS.  Byte-identical layout. */
#define DSPI_WIRE_MAX_CHANNELS 11
#define DSPI_WIRE_MAX_OUTPUTS  9
#define DSPI_WIRE_FORMAT_VERSION 6
#define DSPI_PLATFORM_RP2040 0
#define DSPI_PLATFORM_RP2350 1
typedef struct __attribute__((packed)) {
    struct __attribute__((packed)) { uint8_t format_version, platform_id, num_channels, num_output_channels, num_input_channels, max_bands;
                                     uint16_t payload_length, fw_version_major, fw_version_minor; uint32_t reserved; } header;          /*  16 */
    struct __attribute__((packed)) { float preamp_gain_db; uint8_t bypass, loudness_enabled, reserved[2];
                                     float loudness_ref_spl, loudness_intensity_pct; } global;                                        /*  16 */
    struct __attribute__((packed)) { uint8_t enabled, preset, itd_enabled, reserved; float custom_fc, custom_feed_db; uint32_t reserved2; } crossfeed;   /* 16 */
    struct __attribute__((packed)) { float gain_db[3]; uint8_t mute[3], reserved; } legacy;                                           /*  16 */
    struct __attribute__((packed)) { float delay_ms[DSPI_WIRE_MAX_CHANNELS]; } delays;                                                /*  44 */
    struct __attribute__((packed)) { uint8_t enabled, phase_invert, reserved[2]; float gain_db; } crosspoints[2][DSPI_WIRE_MAX_OUTPUTS];   /* 144 */
    struct __attribute__((packed)) { uint8_t enabled, mute, reserved[2]; float gain_db, delay_ms; } outputs[DSPI_WIRE_MAX_OUTPUTS];   /* 108 */
    struct __attribute__((packed)) { uint8_t num_pin_outputs, pins[5], reserved[2]; } pins;                                           /*   8 */
    struct __attribute__((packed)) { uint8_t type, reserved[3]; float freq, q, gain_db; } eq[DSPI_WIRE_MAX_CHANNELS][DSPI_MAX_BANDS]; /* 2112 */
    char channel_names[DSPI_WIRE_MAX_CHANNELS][32];                                                                                    /* 352 */
    struct __attribute__((packed)) { uint8_t output_types[4], bck_pin, mck_pin, mck_enabled, mck_multiplier, reserved[8]; } i2s_config;    /* 16 */
    struct __attribute__((packed)) { uint8_t enabled, speed, lookahead, reserved; float amount, max_gain_db, gate_threshold_db; } leveller; /* 16 */
    struct __attribute__((packed)) { float preamp_db[2]; uint8_t reserved[8]; } preamp;                                               /*  16 */
    struct __attribute__((packed)) { float master_volume_db; uint8_t reserved[12]; } master_volume;                                   /*  16 */
} dspi_wire_bulk_params;
#ifdef __cplusplus
static_assert(sizeof(dspi_wire_bulk_params) == 2896, "WireBulkParams");
#else
_Static_assert(sizeof(dspi_wire_bulk_params) == 2896, "WireBulkParams");
#endif

/* The globals bulk_params_apply() writes that feed the DSP path (bulk_params.c:22-43), for one device. */
typedef struct {
    int32_t platform;                        /* DSPI_PLATFORM_*: 11 channels / 9 outputs or 7 / 5                  */
    float preamp_db[2], preamp_linear[2]; int32_t preamp_q28[2];          /* global_preamp_db / _linear / _mul     */
    float master_volume_db, master_volume_linear; int32_t master_volume_q15;
    uint8_t bypass_master_eq, loudness_enabled, reserved0[2];
    float loudness_ref_spl, loudness_intensity_pct;
    dspi_crossfeed_config crossfeed;
    dspi_leveller_config leveller;
    float legacy_gain_db[3], legacy_gain_linear[3]; int32_t legacy_gain_mul[3]; uint8_t legacy_mute[3], reserved1;
    float channel_delays_ms[DSPI_WIRE_MAX_CHANNELS];
    dspi_matrix_crosspoint crosspoints[2][DSPI_WIRE_MAX_OUTPUTS];        /* gain_linear by the firmware's db_to_linear */
    dspi_output_channel outputs[DSPI_WIRE_MAX_OUTPUTS];                  /* delay_samples is filled by dspi_bulk_state_to_chain_* */
    dspi_eq_param recipes[DSPI_WIRE_MAX_CHANNELS][DSPI_MAX_BANDS];       /* filter_recipes[][]                      */
} dspi_bulk_state;

/* power-on values of those globals (usb_audio.c:148-211, leveller.h:69-74, matrix defaults config.h) */
void dspi_bulk_state_defaults(dspi_bulk_state *st, int platform);
/* bulk_params_apply(), bulk_params.c:178-377, on `st` instead of the firmware's globals.  Same return codes:
 * 0 ok, -1 format version, -2 platform, -3 channel counts, -4 payload length.  Gains go through the
 * firmware's own db_to_linear (a 4-term Taylor series clamped to [-60, +20] dB, bulk_params.c:49-56;
 * SURVEY quirk 2) unless exact_db != 0 (then 10^(dB/20)); master volume always uses powf (:361-374). */
int dspi_bulk_params_apply(const dspi_wire_bulk_params *in, dspi_bulk_state *st, int exact_db);
/* bulk_params_collect(), bulk_params.c:62-172 (pins, names and I2S sections are control plane: zero) */
void dspi_bulk_params_collect(const dspi_bulk_state *st, dspi_wire_bulk_params *out);
/* What the main loop derives after a successful apply (main.c:1137-1139 and the pending-flag handlers):
 * dsp_recalculate_all_filters, dsp_update_delay_samples, loudness table + row for the host volume,
 * crossfeed and leveller coefficients.  `previous` (may be NULL) supplies the filters whose state must
 * survive (dsp_compute_coefficients only clears state on a topology flip). */
int dspi_bulk_state_to_chain_f32(const dspi_bulk_state *st, float sample_rate, int16_t host_volume_8_8, int host_mute,
                                 dspi_chain_params_f32 *params, dspi_biquad_f32 biquads[11][DSPI_MAX_BANDS]);
int dspi_bulk_state_to_chain_q28(const dspi_bulk_state *st, float sample_rate, int16_t host_volume_8_8, int host_mute,
                                 dspi_chain_params_q28 *params, dspi_biquad_q28 biquads[7][DSPI_MAX_BANDS]);

/* ---- preset slot images: PresetSlot v12, flash_storage.c:139-189 ------------ */
/* One flash sector per slot: 12-byte header (magic "DSP3", data version, slot index, CRC-32 of everything
 * after the header) + the packed DSP state.  Device preset dumps load directly into a dspi_bulk_state and
 * from there (dspi_bulk_state_to_chain_*) into a chain engine. */
#define DSPI_PRESET_SLOT_MAGIC   0x44535033u     /* flash_storage.c:67 */
#define DSPI_PRESET_SLOT_VERSION 12              /* :71 */
#define DSPI_PRESET_OK       0                   /* config.h:262-266 */
#define DSPI_PRESET_ERR_CRC  3
size_t dspi_preset_slot_size(int platform);                       /* sizeof(PresetSlot) on that platform */
uint32_t dspi_crc32(const void *data, size_t len);               /* flash_storage.c:282-291 (reflected 0xEDB88320) */
/* validate_slot() (:750-760) + apply_slot_to_live() (:597-744) + apply_master_volume_from_mode() (:580-590),
 * i.e. the state part of preset_load(): DSPI_PRESET_OK, or DSPI_PRESET_ERR_CRC when magic, slot index or CRC
 * do not match (state untouched).  Gains use flash_storage.c's db_to_linear (powf, :302-306), not the Taylor
 * series of the bulk path.  master_volume_mode / dir_master_volume_db are the directory's settings
 * (MASTER_VOLUME_MODE_INDEPENDENT = 0: use dir_master_volume_db; 1: the slot's own value when version >= 12). */
int dspi_preset_slot_apply(const void *slot, size_t len, uint8_t slot_index, uint8_t master_volume_mode, float dir_master_volume_db,
                           dspi_bulk_state *st);
/* collect_live_state() (:464-556): writes dspi_preset_slot_size(st->platform) bytes with a valid header and CRC.
 * dspi_bulk_state carries the DSP state only: output pins, channel names, output types and the I2S fields are written as
 * ZERO, so the image is for exchanging DSP state between hosts of this library (apply ignores those fields) - do NOT
 * flash it onto a device, which would load the zeros over its pin / name / I2S configuration. */
int dspi_preset_slot_collect(const dspi_bulk_state *st, uint8_t slot_index, void *out, size_t cap);

/* ---- S/PDIF (IEC 60958) subframe encoder: the step after the chain -------------------------- */
/* What stereo_to_spdif_producer_give_s32() does with every S/PDIF producer buffer
 * (pico_audio_spdif_multi/sample_encoding.cpp:42-50 -> spdif_update_subframe,
 * include/pico/audio_spdif/sample_encoding.h:27-50), together with the preamble / channel-status /
 * validity-user-status-parity stamping of init_spdif_buffer (audio_spdif.c:99-114) and the
 * block-position fix-up at DMA start (:372-388): 24-bit words in, 64-bit biphase-mark subframes out,
 * ready for the 2-bits-per-cell PIO serialiser. */
typedef struct { uint32_t l, h; } dspi_spdif_subframe;           /* spdif_subframe_t, sample_encoding.h:20-23 */
/* the reference's 256-entry table (audio_spdif.c:141-153), for hosts that keep its table-driven encoder */
void dspi_spdif_lookup_table(uint32_t table[256]);
/* words: [n_streams][frames][2] int32 (bits 23:0 used) - the layout dspi_chain_process_* writes with
 * n_streams = 4 * n_instances; subframes: [n_streams][frames][2] {l, h}.  Frame n of every stream sits at
 * block position (block_pos0 + n) % 192; channel_status = the 5 consumer status bytes (audio_spdif.c:82-88).
 * _device is asynchronous on `cuda_stream` (a cudaStream_t, may be NULL); _host copies in and out. */
int dspi_spdif_encode_device(int device, const int32_t *d_words, uint64_t n_streams, uint32_t frames, uint32_t block_pos0,
                             const uint8_t channel_status[5], dspi_spdif_subframe *d_subframes, void *cuda_stream);
int dspi_spdif_encode_host(int device, const int32_t *words, uint64_t n_streams, uint32_t frames, uint32_t block_pos0,
                           const uint8_t channel_status[5], dspi_spdif_subframe *subframes);

/* pinned host memory helpers */
void *dspi_host_alloc(size_t bytes);
void dspi_host_free(void *p);
/* Bind the calling thread - and the memory it allocates from now on - to the NUMA node of the device's PCIe link
 * (sysfs numa_node + sched_setaffinity + set_mempolicy), so that staging memory allocated afterwards is local to the
 * link.  Returns the node, or -1 when the topology is not exposed (nothing changed).  Call before dspi_host_alloc. */
int dspi_bind_host_to_device(int device);

#ifdef __cplusplus
}
#endif
#endif /* DSPI_B200_H */
