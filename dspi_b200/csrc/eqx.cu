// eqx.cu — the EQ engine over several GPUs of one box, single process (SURVEY.md §8b `devices[], n_devices`, §8e).
//
// Channels shard into contiguous ranges, one per device; coefficients, filter state and staging buffers of a range live
// on its owner and never move.  There is no cross-channel dependency, so nothing is exchanged between devices: the group
// is N independent dspi_eq engines behind one handle, and a sample block [C][T] - in pinned host memory, or resident on
// the first device of the group ("root") - is fanned out by range:
//   host block  : every device runs its own staged PCIe pipeline (engine.cu eq_process_remote_*) concurrently
//   root block  : the root processes its range in place; every other device pulls its rows over NVLink (peer access),
//                 processes them in its staging buffers and pushes them back - the same three-stream pipeline with the
//                 root's memory in the role of the host, which keeps both directions of each link busy.
// Results are bit-identical to one engine over all channels (tests/test_eqx_gpu.py): sharding changes no bit.
// The multi-PROCESS form of the same thing (one rank per GPU, NCCL send/recv) is dspi_b200/sharding.py.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "eq_kernels.cuh"

namespace {
int failx(int code, const char *fmt, ...)
{
    size_t cap = 0;
    char *buf = dspi::error_buffer(&cap);
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, cap, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace

struct dspi_eqx {
    dspi_eqx_desc desc;
    uint32_t n;                       // devices
    std::vector<dspi_eq *> eng;
    std::vector<uint32_t> lo, hi;     // channel range of each device
    std::vector<int> peer_ok;         // device k can reach the root's memory
};

extern "C" {

int dspi_eqx_shard_range(uint32_t n_channels, uint32_t n_devices, uint32_t k, uint32_t *lo, uint32_t *hi)
{
    if (n_devices == 0 || k >= n_devices || !lo || !hi) return failx(DSPI_EINVAL, "bad shard query");
    // contiguous ranges on 64-channel boundaries (one warp group of K1 never straddles two devices), sizes within one unit
    const uint32_t unit = 64;
    const uint64_t units = ((uint64_t)n_channels + unit - 1) / unit;
    uint64_t a = units * k / n_devices * unit, b = units * (k + 1) / n_devices * unit;
    if (a > n_channels) a = n_channels;
    if (b > n_channels) b = n_channels;
    *lo = (uint32_t)a;
    *hi = (uint32_t)b;
    return DSPI_OK;
}

int dspi_eqx_destroy(dspi_eqx *x)
{
    if (!x) return DSPI_OK;
    for (dspi_eq *e : x->eng) if (e) dspi_eq_destroy(e);
    delete x;
    return DSPI_OK;
}

int dspi_eqx_create(dspi_eqx **out, const dspi_eqx_desc *desc)
{
    if (!out || !desc) return failx(DSPI_EINVAL, "null argument");
    *out = nullptr;
    if (desc->n_devices == 0 || desc->n_devices > DSPI_MAX_DEVICES) return failx(DSPI_EINVAL, "n_devices must be 1..%d", DSPI_MAX_DEVICES);
    if (desc->n_channels == 0) return failx(DSPI_EINVAL, "n_channels must be > 0");
    for (uint32_t a = 0; a < desc->n_devices; a++)
        for (uint32_t b = a + 1; b < desc->n_devices; b++)
            if (desc->devices[a] == desc->devices[b]) return failx(DSPI_EINVAL, "device %d listed twice", desc->devices[a]);
    dspi_eqx *x = new (std::nothrow) dspi_eqx();
    if (!x) return failx(DSPI_ENOMEM, "host allocation failed");
    x->desc = *desc;
    x->n = desc->n_devices;
    x->eng.assign(x->n, nullptr);
    x->lo.resize(x->n); x->hi.resize(x->n); x->peer_ok.assign(x->n, 0);
    for (uint32_t k = 0; k < x->n; k++) {
        dspi_eqx_shard_range(desc->n_channels, x->n, k, &x->lo[k], &x->hi[k]);
        if (x->hi[k] == x->lo[k]) continue;                                  // more devices than 64-channel units
        dspi_eq_desc d;
        memset(&d, 0, sizeof(d));
        d.arith = desc->arith; d.n_channels = x->hi[k] - x->lo[k]; d.n_bands = desc->n_bands; d.device = desc->devices[k];
        const int rc = dspi_eq_create(&x->eng[k], &d);
        if (rc) { dspi_eqx_destroy(x); return rc; }
    }
    // peer access towards the root (device 0 of the list): needed only by dspi_eqx_process_root
    for (uint32_t k = 1; k < x->n; k++) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, desc->devices[k], desc->devices[0]) == cudaSuccess && can) {
            cudaSetDevice(desc->devices[k]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(desc->devices[0], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) x->peer_ok[k] = 1;
        }
        cudaGetLastError();
    }
    *out = x;
    return DSPI_OK;
}

static int route(dspi_eqx *x, uint32_t ch0, uint32_t n, void *rows, size_t row_bytes, bool upload)
{
    if (!x || !rows) return failx(DSPI_EINVAL, "null argument");
    if ((uint64_t)ch0 + n > x->desc.n_channels) return failx(DSPI_ERANGE, "channels [%u, %u) outside group of %u", ch0, ch0 + n, x->desc.n_channels);
    for (uint32_t k = 0; k < x->n; k++) {
        const uint32_t a = ch0 > x->lo[k] ? ch0 : x->lo[k], b = (ch0 + n) < x->hi[k] ? (ch0 + n) : x->hi[k];
        if (a >= b) continue;
        char *p = (char *)rows + (size_t)(a - ch0) * row_bytes;
        const int rc = upload ? dspi_eq_upload_biquads(x->eng[k], a - x->lo[k], b - a, p) : dspi_eq_download_biquads(x->eng[k], a - x->lo[k], b - a, p);
        if (rc) return rc;
    }
    return DSPI_OK;
}

int dspi_eqx_upload_biquads(dspi_eqx *x, uint32_t ch0, uint32_t n, const void *biquads)
{
    if (!x) return failx(DSPI_EINVAL, "null argument");
    const size_t row = (size_t)DSPI_MAX_BANDS * (x->desc.arith == DSPI_ARITH_Q28 ? sizeof(dspi_biquad_q28) : sizeof(dspi_biquad_f32));
    return route(x, ch0, n, const_cast<void *>(biquads), row, true);
}

int dspi_eqx_download_biquads(dspi_eqx *x, uint32_t ch0, uint32_t n, void *biquads)
{
    if (!x) return failx(DSPI_EINVAL, "null argument");
    const size_t row = (size_t)DSPI_MAX_BANDS * (x->desc.arith == DSPI_ARITH_Q28 ? sizeof(dspi_biquad_q28) : sizeof(dspi_biquad_f32));
    return route(x, ch0, n, biquads, row, false);
}

int dspi_eqx_process_host(dspi_eqx *x, void *h_samples, uint32_t T)
{
    if (!x || !h_samples) return failx(DSPI_EINVAL, "null argument");
    if (T == 0) return DSPI_OK;
    int rc = DSPI_OK;
    for (uint32_t k = 0; k < x->n && rc == DSPI_OK; k++)                     // every device's pipeline is enqueued before any is awaited
        if (x->eng[k]) rc = dspi::eq_process_remote_enqueue(x->eng[k], (char *)h_samples + (size_t)x->lo[k] * T * 4, T, 0, x->hi[k] - x->lo[k]);
    for (uint32_t k = 0; k < x->n; k++)
        if (x->eng[k]) { const int w = dspi::eq_process_remote_wait(x->eng[k]); if (rc == DSPI_OK) rc = w; }
    return rc;
}

int dspi_eqx_process_root(dspi_eqx *x, void *d_samples_on_root, uint32_t T, uint32_t ld)
{
    if (!x || !d_samples_on_root) return failx(DSPI_EINVAL, "null argument");
    if (T == 0) return DSPI_OK;
    if (ld != T && x->n > 1) return failx(DSPI_EINVAL, "the root block must be dense (ld == T) when more than one device takes part");
    for (uint32_t k = 1; k < x->n; k++)
        if (x->eng[k] && !x->peer_ok[k]) return failx(DSPI_ENODEV, "device %d cannot access the memory of root device %d", x->desc.devices[k], x->desc.devices[0]);
    int rc = DSPI_OK;
    for (uint32_t k = 1; k < x->n && rc == DSPI_OK; k++)                     // the peers pull / process / push over NVLink ...
        if (x->eng[k]) rc = dspi::eq_process_remote_enqueue(x->eng[k], (char *)d_samples_on_root + (size_t)x->lo[k] * T * 4, T, 0, x->hi[k] - x->lo[k]);
    if (rc == DSPI_OK && x->eng[0])                                          // ... while the root works on its own rows in place
        rc = dspi_eq_process_device_range(x->eng[0], d_samples_on_root, T, ld, 0, x->hi[0] - x->lo[0]);
    for (uint32_t k = 1; k < x->n; k++)
        if (x->eng[k]) { const int w = dspi::eq_process_remote_wait(x->eng[k]); if (rc == DSPI_OK) rc = w; }
    if (x->eng[0]) { const int w = dspi_eq_sync(x->eng[0]); if (rc == DSPI_OK) rc = w; }
    return rc;
}

uint64_t dspi_eqx_launch_count(dspi_eqx *x)
{
    uint64_t n = 0;
    if (x) for (dspi_eq *e : x->eng) if (e) n += dspi_eq_launch_count(e);
    return n;
}

}  // extern "C"
