// eq_f32_kernel.cuh — body of K1 (see eq_f32.cu for the design notes).  Kept in a header that is safe
// for runtime compilation: the ahead-of-time build instantiates it for arbitrary topologies, the
// runtime compiler (eq_jit.cu) for ONE topology vector as a template constant.
#pragma once
#include "eq_core.cuh"

namespace dspi {
namespace k1 {

using namespace core;

constexpr int kTileT = 32;          // samples per shared-memory tile row: 128 B == swizzle span
constexpr int kStages = 3;

// SIG::enabled: the launch is known to have ONE topology vector SIG::word for every channel
struct NoSig { static constexpr bool enabled = false; static constexpr unsigned long long word = 0; };
template <unsigned long long W> struct SigWord { static constexpr bool enabled = true; static constexpr unsigned long long word = W; };

template <typename V, bool FUSED, int NB, bool DYN, typename SIG>
__device__ __forceinline__ void eq_f32_body(const CUtensorMap &tmap, float *__restrict__ samples, uint32_t ld, V *__restrict__ coef,
              const uint64_t *__restrict__ modes, uint32_t n_groups, uint32_t n_rows, uint32_t T, uint32_t nb_active, uint32_t use_tma, uint32_t dbg, unsigned long long nz_bits, uint32_t slice_tiles, uint32_t *__restrict__ sched)
{
    constexpr int CPL = Lanes<V>::CPL;
    constexpr int kRows = 32 * CPL;
    constexpr int kWarps = 16 / CPL;
    constexpr uint32_t kStageBytes = kRows * kTileT * 4;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[16 / Lanes<V>::CPL][kStages];

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
        // register-tile (straight-line) path: all-biquad warps in the ahead-of-time kernels; warps whose
        // channels all carry the compiled-in topology vector in a runtime-specialised kernel
        bool straight = bank.all_tdf2;
        if constexpr (SIG::enabled) straight = bank.template sig_match<SIG::word>();

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
            if (straight && tile_valid == kTileT && !(dbg & 4u)) {
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
                    if (!(dbg & 1u)) {                                 // DSPI_DBG=1: data path only
                        if constexpr (SIG::enabled) bank.template run_sig<SIG::word>(x, nz);
                        else bank.run(x, kSub, nz);
                    }
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


}  // namespace k1
}  // namespace dspi
