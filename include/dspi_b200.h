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
int dspi_eq_process_host(dspi_eq *e, void *h_samples, uint32_t T);
int dspi_eq_sync(dspi_eq *e);
/* cudaStream_t of the engine (so callers can order their own work / events) */
void *dspi_eq_stream(dspi_eq *e);
/* number of kernel launches issued by this engine so far */
uint64_t dspi_eq_launch_count(dspi_eq *e);

/* pinned host memory helpers */
void *dspi_host_alloc(size_t bytes);
void dspi_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* DSPI_B200_H */
