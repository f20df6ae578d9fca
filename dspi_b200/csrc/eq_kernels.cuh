// eq_kernels.cuh — declarations shared by the EQ kernels and the engine (C-ABI) code.
#pragma once
#include <utility>
#include <vector>

#include "eq_modes.cuh"
#include "dspi_b200.h"

namespace dspi {

static_assert(kMaxBands == DSPI_MAX_BANDS, "band stride");

struct EqLaunch {
    CUtensorMap tmap;      // [C rows][T] tiled map, box {32, rows-per-warp}, SWIZZLE_128B
    void *samples;         // device base (fallback path)
    uint32_t ld;           // row stride in elements
    void *coef;            // packed coefficient/state store
    const uint64_t *modes; // f32 only
    uint32_t n_groups;     // groups of (32 * channels-per-lane) channels
    uint32_t n_rows;       // valid channel rows in `samples`
    uint32_t T;
    uint32_t n_bands;
    uint32_t use_tma;
    uint32_t dbg;          // diagnostics only (env DSPI_DBG): 1 = skip arithmetic, 2 = skip HBM traffic, 4 = force column path, 8 = dynamic time-slice schedule
    uint32_t *sched;       // K1 dynamic scheduler words: [0] item counter, [1 + g] slices completed by group g (or nullptr)
    int n_sms;             // SM count of the device
};

// cudaFuncSetAttribute is per device: remembers, per kernel (one flag word per call site), which devices are done
struct PerDeviceOnce {
    unsigned long long done = 0;
    bool needs(int *device_out)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        *device_out = dev;
        return dev < 0 || dev >= 64 || !((done >> dev) & 1ull);
    }
    void mark(int dev) { if (dev >= 0 && dev < 64) done |= 1ull << dev; }
};

// per-thread message buffer behind dspi_last_error() (engine.cu)
char *error_buffer(size_t *cap);

// K1 — float cascade.  cpl: channels per lane (1 scalar FFMA, 2 packed FFMA2)
cudaError_t launch_eq_f32(const EqLaunch &a, bool fused, int cpl, cudaStream_t stream);
cudaError_t launch_pack_f32(const dspi_biquad_f32 *aos, uint32_t ch0, uint32_t n, float *coef, uint64_t *modes, int cpl, cudaStream_t stream);
cudaError_t launch_unpack_f32(dspi_biquad_f32 *aos, uint32_t ch0, uint32_t n, const float *coef, int cpl, cudaStream_t stream);

// K2 — Q28 cascade (1 channel per lane)
cudaError_t launch_eq_q28(const EqLaunch &a, cudaStream_t stream);
cudaError_t launch_pack_q28(const dspi_biquad_q28 *aos, uint32_t ch0, uint32_t n, int32_t *coef, cudaStream_t stream);
cudaError_t launch_unpack_q28(dspi_biquad_q28 *aos, uint32_t ch0, uint32_t n, const int32_t *coef, cudaStream_t stream);

}  // namespace dspi

// ---- engine-internal interface (engine.cu), used by the chain engines which run their EQ rows
// through K1: same packed store, same kernels, on a stream of the caller's choosing -------------
struct dspi_eq;
namespace dspi {
void *eq_aos_mirror(dspi_eq *e);                                          // device Biquad[c_pad][12], reference layout
int eq_pack_range(dspi_eq *e, uint32_t ch0, uint32_t n, cudaStream_t s);   // mirror -> packed store (coefficients and state), synchronous
int eq_unpack_range(dspi_eq *e, uint32_t ch0, uint32_t n, cudaStream_t s); // packed store -> mirror (state), asynchronous on s
// rows with skip[ch] != 0 keep their whole cascade frozen (all bands treated as bypassed, state untouched):
// usb_audio.c:721-728 (bypass_master_eq), :879-884 (muted / disabled outputs).  `d_skip` is [n_channels]
// device memory owned by the caller; call again after changing it.
int eq_set_skip(dspi_eq *e, const uint8_t *d_skip, cudaStream_t s);
int eq_process_on(dspi_eq *e, void *d_samples, uint32_t T, uint32_t ld, cudaStream_t s);
int eq_process_range_on(dspi_eq *e, void *d_rows, uint32_t T, uint32_t ld, uint32_t ch0, uint32_t n, cudaStream_t s);
// staged copy-in / kernel / copy-out pipeline of channels [ch0, ch0 + n_ch) against `remote` rows [n_ch][T] (pinned host
// memory or a peer GPU's memory); enqueue does not block, wait returns when the results are back in `remote`
int eq_process_remote_enqueue(dspi_eq *e, void *remote, uint32_t T, uint32_t ch0, uint32_t n_ch);
int eq_process_remote_wait(dspi_eq *e);
void eq_state_sections(dspi_eq *e, std::vector<std::pair<void *, size_t>> &out);   // coefficient + state store (and topology words)
int eq_state_imported(dspi_eq *e, cudaStream_t s);
// coeff.cu: dsp_compute_coefficients() for channels [ch0, ch0 + n) of a mirror, recipes [n][12] on the device (clamped in place)
cudaError_t launch_coeffs(bool q28, dspi_eq_param *d_recipes, void *d_aos, uint32_t ch0, uint32_t n, float fs, cudaStream_t stream);
cudaError_t launch_skip_q28(int32_t *coef, const uint8_t *skip, uint32_t n, cudaStream_t stream);
cudaError_t launch_mask_modes(const uint64_t *raw, const uint8_t *skip, uint64_t *eff, uint32_t n, cudaStream_t stream);
}  // namespace dspi
