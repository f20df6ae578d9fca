// eq_f32.cu — K1: float32 10-band EQ cascade (Cytomic SVF / TDF2 biquad hybrid) for
// thousands of independent channels, sm_100a.
//
// Reference semantics: dsp_process_channel_block(), firmware/DSPi/dsp_pipeline.c:281-365
// (band outer / sample inner, in place; the per-sample twin :256-279 yields the same values).
//
// Mapping
//   * one warp owns a GROUP of 32*CPL channels for the whole launch; lane L carries channel
//     L (and L+32 when CPL==2, packed in the two halves of an f32x2 register pair) — the
//     serial sample recurrence and all coefficients/state of the 10 bands live in registers;
//   * samples are channel-major [C][T] in HBM; each warp streams its [32*CPL][32] tiles
//     through a private 3-stage shared-memory ring with TMA (cp.async.bulk.tensor, 128-byte
//     swizzle, mbarrier completion) and writes results back with TMA stores from the same
//     buffers — no block-wide synchronisation anywhere;
//   * CPL==2 uses Blackwell's packed FFMA2/FMUL2/FADD2 (fma/mul/add.rn.ftz.f32x2): the FMA
//     pipe sees the same number of lane-operations but only half the issue slots, which
//     leaves room for LDS/STS/branches next to a saturated FMA pipe.
//
// Arithmetic is written with explicit-rounding intrinsics only, so nvcc can neither contract
// nor reassociate: FUSED follows GCC's -ffp-contract=fast pattern (what arm-none-eabi-gcc
// emits for the RP2350), !FUSED rounds every operation separately.  -ftz=true gives the
// firmware's FZ mode (main.c:593-600).  Negations are kept out of the inner loops (packed
// ops have no free negate): a1/a2 are stored negated, `a - b` is fma(b, -1, a) (exact), and
// the SVF state alternates sign every sample (see svf_tile() in eq_core.cuh).
#include "eq_core.cuh"

namespace dspi {
namespace {

using namespace core;

constexpr int kTileT = 32;          // samples per shared-memory tile row: 128 B == swizzle span
constexpr int kStages = 3;

template <typename V, bool FUSED, int NB, bool DYN>
__global__ void __launch_bounds__(256 * 2 / Lanes<V>::CPL, 1)
eq_f32_kernel(const __grid_constant__ CUtensorMap tmap, float *__restrict__ samples, uint32_t ld, V *__restrict__ coef,
              const uint64_t *__restrict__ modes, uint32_t n_groups, uint32_t n_rows, uint32_t T, uint32_t nb_active, uint32_t use_tma, uint32_t dbg, unsigned long long nz_bits, uint32_t slice_tiles, uint32_t *__restrict__ sched)
{
    constexpr int CPL = Lanes<V>::CPL;
    constexpr int kRows = 32 * CPL;
    constexpr int kWarps = 16 / CPL;
    constexpr uint32_t kStageBytes = kRows * kTileT * 4;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[kWarps][kStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    uint8_t *my_smem = smem_raw + (size_t)warp * kStages * kStageBytes;
    uint64_t *full = bars[warp];
    if (lane == 0) {
        if (use_tma) prefetch_tmap(&tmap);
        for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const V nz = v_bits<V>(nz_bits);                            // (-0.0, -0.0): see mulx()
    const uint32_t ntiles = (T + kTileT - 1) / kTileT;
    const bool mem_on = !(dbg & 2u);                            // diagnostics: DSPI_DBG=2 runs the arithmetic without HBM traffic
    const uint32_t sw = (lane & 7) << 4;                        // 128B-swizzle XOR for this lane's rows

    // ---- work distribution -------------------------------------------------------------------
    // A work item is (group g, time slice k): `slice_tiles` consecutive tiles of the 32*CPL channels
    // of group g.  Static mode (sched == nullptr): one item per warp = the whole launch of group
    // blockIdx.x * kWarps + warp.  Dynamic mode: a persistent grid pulls items from an atomic counter
    // in slice-major order; the filter state of a group travels from slice to slice through the
    // coefficient store, ordered by a per-group completion counter (release/acquire).  65536
    // channels are 1024 groups for 592 warp schedulers: no static split can load them evenly, the
    // time slices can.
    const uint32_t n_slices = DYN ? (ntiles + slice_tiles - 1) / slice_tiles : 1;
    const uint32_t n_items = DYN ? n_groups * n_slices : 0;
    uint32_t tcount = 0;                                        // tiles this warp has pushed through its ring (stage / parity bookkeeping)
    for (;;) {
        uint32_t g, tile_begin, tile_end, slice = 0;
        if constexpr (DYN) {
            uint32_t item = 0;
            if (lane == 0) item = atomicAdd(&sched[0], 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) break;
            slice = item / n_groups;
            g = item - slice * n_groups;
            tile_begin = slice * slice_tiles;
            tile_end = min(ntiles, tile_begin + slice_tiles);
            if (slice > 0) {                                    // wait until the previous slice of this group has published its state
                if (lane == 0) {
                    const volatile uint32_t *flag = sched + 1 + g;
                    while (*flag < slice) __nanosleep(100);
                    __threadfence();
                }
                __syncwarp();
            }
        } else {
            g = blockIdx.x * kWarps + warp;
            if (g >= n_groups) break;                           // warps are fully independent
            tile_begin = 0;
            tile_end = ntiles;
        }
        const int c0 = g * kRows;                               // first channel (row) of this group

        auto issue_load = [&](uint32_t tile, uint32_t seq) {    // lane 0 only; seq = position in this warp's ring sequence
            const uint32_t s = seq % kStages;
            mbar_arrive_expect_tx(&full[s], kStageBytes);
            tma_load_2d(my_smem + s * kStageBytes, &tmap, &full[s], tile * kTileT, c0);
        };
        if (use_tma && mem_on && lane == 0) {
            if constexpr (DYN) tma_store_wait_read<0>();        // ring buffers of the previous item are drained
            for (uint32_t j = 0; j + 1 < kStages && tile_begin + j < tile_end; j++) issue_load(tile_begin + j, tcount + j);
        }

        // ---- coefficients, state and topology of every band -> registers ----------------------
        EqBank<V, FUSED, NB> bank;
        V *my_coef = coef + (size_t)g * kMaxBands * 8 * 32 + lane;
        {
            const uint64_t *mp[CPL];
#pragma unroll
            for (int h = 0; h < CPL; h++) mp[h] = modes + (size_t)g * kRows + h * 32 + lane;
            bank.load(my_coef, mp, nb_active, DYN);
        }

        // ---- stream the tiles of this item --------------------------------------------------------
        for (uint32_t tile = tile_begin; tile < tile_end; tile++, tcount++) {
            const uint32_t s = tcount % kStages;
            uint8_t *buf = my_smem + s * kStageBytes;
            if (use_tma) {
                if (mem_on) mbar_wait(&full[s], (tcount / kStages) & 1);
            } else {                                                // plain-load fallback (odd strides / unaligned bases)
                const uint32_t t = tile * kTileT + lane;
                for (int r = 0; r < kRows; r++) {
                    const uint32_t ch = c0 + r;
                    float v = 0.0f;
                    if (t < T && ch < n_rows) v = samples[(size_t)ch * ld + t];
                    *reinterpret_cast<float *>(buf + r * 128 + ((((lane >> 2) << 4) ^ ((r & 7) << 4)) | ((lane & 3) << 2))) = v;
                }
                __syncwarp();
            }

            const int tile_valid = min((int)kTileT, (int)(T - tile * kTileT));
            if (bank.all_tdf2 && tile_valid == kTileT && !(dbg & 4u)) {
                // ---- all-biquad warps: register tiles of kSub samples, straight-line over the 10 bands ----
    #pragma unroll 1
                for (int sub = 0; sub < kTileT / kSub; sub++) {
                    // two 16-byte chunks per row per sub-tile; chunk index XOR (row & 7)
                    V x[kSub];
                    float4 q[CPL][2];
    #pragma unroll
                    for (int h = 0; h < CPL; h++) {
                        const uint8_t *row = buf + (lane + 32 * h) * 128;
                        q[h][0] = *reinterpret_cast<const float4 *>(row + (((2 * sub) << 4) ^ sw));
                        q[h][1] = *reinterpret_cast<const float4 *>(row + (((2 * sub + 1) << 4) ^ sw));
                    }
    #pragma unroll
                    for (int i = 0; i < kSub; i++) {
                        float part[CPL];
    #pragma unroll
                        for (int h = 0; h < CPL; h++) {
                            const float4 &qq = q[h][i >> 2];
                            part[h] = (i & 3) == 0 ? qq.x : (i & 3) == 1 ? qq.y : (i & 3) == 2 ? qq.z : qq.w;
                        }
                        v_make(x[i], part);
                    }
                    if (!(dbg & 1u)) bank.run(x, kSub, nz);            // DSPI_DBG=1: data path only
    #pragma unroll
                    for (int h = 0; h < CPL; h++) {
                        uint8_t *row = buf + (lane + 32 * h) * 128;
                        *reinterpret_cast<float4 *>(row + (((2 * sub) << 4) ^ sw)) =
                            make_float4(Lanes<V>::get(x[0], h), Lanes<V>::get(x[1], h), Lanes<V>::get(x[2], h), Lanes<V>::get(x[3], h));
                        *reinterpret_cast<float4 *>(row + (((2 * sub + 1) << 4) ^ sw)) =
                            make_float4(Lanes<V>::get(x[4], h), Lanes<V>::get(x[5], h), Lanes<V>::get(x[6], h), Lanes<V>::get(x[7], h));
                    }
                }
            } else {
                // ---- any other topology: band-outer over the tile, re-laid out in place as lane-private
                //      columns of CPL-vectors (sample n of this lane at col[n * 32]) ----
                float4 q[CPL][8];
    #pragma unroll
                for (int h = 0; h < CPL; h++) {
                    const uint8_t *row = buf + (lane + 32 * h) * 128;
    #pragma unroll
                    for (int k = 0; k < 8; k++) q[h][k] = *reinterpret_cast<const float4 *>(row + ((k << 4) ^ sw));
                }
                __syncwarp();                                       // every row is in registers before columns overwrite them
                V *col = reinterpret_cast<V *>(buf) + lane;
    #pragma unroll
                for (int n = 0; n < kTileT; n++) {
                    float part[CPL];
    #pragma unroll
                    for (int h = 0; h < CPL; h++) {
                        const float4 &qq = q[h][n >> 2];
                        part[h] = (n & 3) == 0 ? qq.x : (n & 3) == 1 ? qq.y : (n & 3) == 2 ? qq.z : qq.w;
                    }
                    V v;
                    v_make(v, part);
                    col[n * 32] = v;
                }
                if (!(dbg & 1u)) bank.run_columns(col, tile_valid, nz);
                float back[CPL][kTileT];
    #pragma unroll
                for (int n = 0; n < kTileT; n++) {
                    const V v = col[n * 32];
    #pragma unroll
                    for (int h = 0; h < CPL; h++) back[h][n] = Lanes<V>::get(v, h);
                }
                __syncwarp();
    #pragma unroll
                for (int h = 0; h < CPL; h++) {
                    uint8_t *row = buf + (lane + 32 * h) * 128;
    #pragma unroll
                    for (int k = 0; k < 8; k++)
                        *reinterpret_cast<float4 *>(row + ((k << 4) ^ sw)) = make_float4(back[h][4 * k], back[h][4 * k + 1], back[h][4 * k + 2], back[h][4 * k + 3]);
                }
            }

            if (use_tma) {
                fence_proxy_async_smem();                           // my smem writes -> async proxy
                __syncwarp();
                if (lane == 0 && mem_on) {
                    tma_store_2d(&tmap, buf, tile * kTileT, c0);
                    tma_store_commit();
                    const uint32_t nxt = tile + kStages - 1;        // refill the buffer stored one iteration ago
                    if (nxt < tile_end) {
                        tma_store_wait_read<1>();
                        issue_load(nxt, tcount + kStages - 1);
                    }
                }
            } else {
                __syncwarp();
                const uint32_t t = tile * kTileT + lane;
                for (int r = 0; r < kRows; r++) {
                    const uint32_t ch = c0 + r;
                    const float v = *reinterpret_cast<const float *>(buf + r * 128 + ((((lane >> 2) << 4) ^ ((r & 7) << 4)) | ((lane & 3) << 2)));
                    if (t < T && ch < n_rows) samples[(size_t)ch * ld + t] = v;
                }
                __syncwarp();
            }
        }


        bank.store(my_coef, DYN);                                // filter state back to the coefficient store
        if constexpr (!DYN) break;
        __threadfence();                                        // state visible before the slice counter moves
        __syncwarp();
        if (lane == 0) atomicExch(&sched[1 + g], slice + 1);
    }
    if (use_tma && lane == 0) tma_store_wait_all<0>();          // smem must outlive the bulk reads
}

template <typename V, bool FUSED, int NB>
cudaError_t launch_one(const EqLaunch &a, cudaStream_t stream)
{
    constexpr int CPL = Lanes<V>::CPL;
    constexpr int kWarps = 16 / CPL;
    constexpr size_t smem = (size_t)kWarps * kStages * (32 * CPL) * kTileT * 4;
    auto kern = eq_f32_kernel<V, FUSED, NB, false>;
    auto kern_dyn = eq_f32_kernel<V, FUSED, NB, true>;
    static bool configured = false;                             // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const uint32_t n_groups = a.n_groups;
    uint32_t grid = (n_groups + kWarps - 1) / kWarps;
    uint32_t slice_tiles = 0;
    uint32_t *sched = nullptr;
    const uint32_t ntiles = (a.T + kTileT - 1) / kTileT;
    // The dynamic schedule is opt-in (DSPI_DBG=8): measured on B200 it loses ~6 % to the static one at
    // 65536 channels - a scheduler left with ONE resident warp runs it well below half the two-warp rate,
    // which eats the balance it buys.
    if (a.sched && a.use_tma && ntiles >= 32 && (a.dbg & 8u)) {
        slice_tiles = 16;
        sched = a.sched;
        if (grid > (uint32_t)a.n_sms) grid = a.n_sms;
        else if (grid < (uint32_t)a.n_sms && (uint32_t)a.n_sms * kWarps <= n_groups * ((ntiles + slice_tiles - 1) / slice_tiles)) grid = a.n_sms;
        cudaError_t e = cudaMemsetAsync(sched, 0, (size_t)(1 + n_groups) * sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
    }
    if (sched)
        kern_dyn<<<grid, kWarps * 32, smem, stream>>>(a.tmap, (float *)a.samples, a.ld, (V *)a.coef, a.modes, n_groups, a.n_rows, a.T, a.n_bands, a.use_tma,
                                                      a.dbg, 0x8000000080000000ull, slice_tiles, sched);
    else
        kern<<<grid, kWarps * 32, smem, stream>>>(a.tmap, (float *)a.samples, a.ld, (V *)a.coef, a.modes, n_groups, a.n_rows, a.T, a.n_bands, a.use_tma,
                                                  a.dbg, 0x8000000080000000ull, 0u, nullptr);
    return cudaGetLastError();
}

template <typename V, bool FUSED>
cudaError_t launch_nb(const EqLaunch &a, cudaStream_t stream)
{
    if (a.n_bands <= 10) return launch_one<V, FUSED, 10>(a, stream);
    return launch_one<V, FUSED, 12>(a, stream);
}

}  // namespace

cudaError_t launch_eq_f32(const EqLaunch &a, bool fused, int cpl, cudaStream_t stream)
{
    if (cpl == 2) return fused ? launch_nb<P2, true>(a, stream) : launch_nb<P2, false>(a, stream);
    return fused ? launch_nb<float, true>(a, stream) : launch_nb<float, false>(a, stream);
}

// ---------------------------------------------------------------------------------------
// Biquad[n][12] (reference layout, AoS) <-> packed device store
// ---------------------------------------------------------------------------------------
__global__ void pack_f32_kernel(const dspi_biquad_f32 *__restrict__ aos, uint32_t ch0, uint32_t n, float *__restrict__ coef,
                                uint64_t *__restrict__ modes, int cpl)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ch = ch0 + i;
    const uint32_t rows = 32 * cpl;
    const uint32_t g = ch / rows, r = ch % rows, lane = r & 31, h = r >> 5;
    uint64_t mw = 0;
    for (int b = 0; b < kMaxBands; b++) {
        const dspi_biquad_f32 &q = aos[(size_t)ch * kMaxBands + b];
        float v[8];
        uint32_t mode;
        if (q.bypass) mode = kModeBypass;
        else if (!q.use_svf) mode = kModeTdf2;
        else mode = q.svf_type == DSPI_FILTER_LOWPASS ? kModeSvfLP : q.svf_type == DSPI_FILTER_HIGHPASS ? kModeSvfHP
                  : q.svf_type == DSPI_FILTER_PEAKING ? kModeSvfPK : kModeSvfSH;
        if (q.use_svf && !q.bypass) {
            v[0] = q.sva1; v[1] = q.sva2; v[2] = q.sva3; v[3] = q.svm0; v[4] = q.svm1; v[5] = q.svm2;
            v[6] = q.svic1eq; v[7] = q.svic2eq;
        } else {
            v[0] = q.b0; v[1] = q.b1; v[2] = q.b2; v[3] = -q.a1; v[4] = -q.a2; v[5] = 0.0f;
            v[6] = q.s1; v[7] = q.s2;
        }
        mw |= (uint64_t)mode << (4 * b);
        for (int k = 0; k < 8; k++) coef[((((size_t)g * kMaxBands + b) * 8 + k) * 32 + lane) * cpl + h] = v[k];
    }
    modes[(size_t)g * rows + h * 32 + lane] = mw;
}

__global__ void unpack_f32_kernel(dspi_biquad_f32 *__restrict__ aos, uint32_t ch0, uint32_t n, const float *__restrict__ coef, int cpl)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ch = ch0 + i;
    const uint32_t rows = 32 * cpl;
    const uint32_t g = ch / rows, r = ch % rows, lane = r & 31, h = r >> 5;
    for (int b = 0; b < kMaxBands; b++) {
        dspi_biquad_f32 &q = aos[(size_t)ch * kMaxBands + b];
        if (q.bypass) continue;
        const float s0 = coef[((((size_t)g * kMaxBands + b) * 8 + 6) * 32 + lane) * cpl + h];
        const float s1 = coef[((((size_t)g * kMaxBands + b) * 8 + 7) * 32 + lane) * cpl + h];
        if (q.use_svf) { q.svic1eq = s0; q.svic2eq = s1; }
        else { q.s1 = s0; q.s2 = s1; }
    }
}

cudaError_t launch_pack_f32(const dspi_biquad_f32 *aos, uint32_t ch0, uint32_t n, float *coef, uint64_t *modes, int cpl, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    pack_f32_kernel<<<(n + 127) / 128, 128, 0, stream>>>(aos, ch0, n, coef, modes, cpl);
    return cudaGetLastError();
}
cudaError_t launch_unpack_f32(dspi_biquad_f32 *aos, uint32_t ch0, uint32_t n, const float *coef, int cpl, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    unpack_f32_kernel<<<(n + 127) / 128, 128, 0, stream>>>(aos, ch0, n, coef, cpl);
    return cudaGetLastError();
}

}  // namespace dspi
