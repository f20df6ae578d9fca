// chain_q28.cu — the whole per-packet signal chain in the RP2040's Q28 fixed-point arithmetic, for
// thousands of independent device instances (2 inputs -> 5 outputs each), bit-exact, sm_100a.
//
// Reference: process_audio_packet(), firmware/DSPi/usb_audio.c:968-1283 (single-core branch
// :1191-1276); fast_mul_q28 dsp_pipeline.c:47-58; fast_mul_q15 config.h:556-567; cascade
// dsp_process_rp2040.S:225-394; crossfeed.c:161-180; leveller.c:275-389; PDM pdm_generator.c:351-397.
//
// Same stage-wise decomposition as chain_f32.cu: pre (unpack, preamp, loudness) -> K2 over the master rows
// -> post (leveller, peaks, crossfeed) -> mix -> K2 over the output rows -> outpost (gain, delay, metering,
// 24-bit words) -> ring update -> modulator, on three streams over packet slices.  The EQ rows run through
// the Q28 cascade kernel of the EQ engine (eq_q28.cu: TMA ring, coefficients pre-split in registers);
// everything is integer-pipe bound (about 27 integer ops per band-sample).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "eq_kernels.cuh"
#include "chain_pdm.cuh"
#include "chain_streams.cuh"
#include "dynamics.cuh"

namespace dspi {
namespace {

constexpr int kOuts = DSPI_CHAINQ_OUTPUTS;
constexpr int kRoles = DSPI_CHAINQ_EQ_CHANNELS;
constexpr int kMaxDelay = DSPI_CHAINQ_MAX_DELAY;
constexpr int kLa = DSPI_LA_SAMPLES;
constexpr int kPkt = DSPI_PACKET_MAX;
constexpr int kXs = 33;                                       // shared-memory column stride: conflict-free for lane = instance AND lane = frame
constexpr int32_t kUnity = 1 << 28;
constexpr int32_t kClipThresh = (1 << 28) + 268;              // config.h:54

enum : uint8_t { F_BYPASS_MASTER = 1, F_LOUD = 2, F_XFEED = 4, F_LEV = 8, F_LOOKAHEAD = 16, F_ANY_DELAY = 32, F_SUB_ON = 64 };
enum : uint8_t { O_ENABLED = 1, O_MUTE = 2, O_PAIR_OFF = 4 };

struct ChainQ {
    uint32_t N, N_pad, nb, max_frames, ldF;        // ldF: row stride of mrow / orow / subq (frames, multiple of 4)
    int32_t *preamp;                               // [2][N_pad]
    uint8_t *flags;
    int32_t *loud_c; int32_t *loud_st; uint8_t *loud_byp;    // [2 j][5][N_pad], [2 side][2 j][2][N_pad], [N_pad]
    int32_t *xf;                                   // [7][N_pad]
    float *lev_c; int32_t *lev_i; float *lev_f; uint32_t *lev_idx; int32_t *lev_la;   // [9][Np], [4][Np] env_l env_r gain gain_prev, [Np] smooth_db, [Np], [2][480][Np]
    int32_t *o_gl, *o_gr, *o_gain; uint8_t *o_flags; int32_t *o_dly;                  // [5][N_pad]
    int32_t *dline; uint32_t *widx_in, *widx_out;  // [5][N_pad][2048], [N_pad]
    int32_t *pdm;                                  // [9][N_pad]
    uint16_t *peaks; uint16_t *clip;               // [7][N_pad], [N_pad]
    int32_t *mrow;                                 // [2 N_pad][ldF] master rows, row = side * N_pad + inst
    int32_t *orow;                                 // [5 N_pad][ldF] output rows, row = o * N_pad + inst
    int32_t *subq;                                 // [N_pad][ldF] Q28 sub samples for the modulator
    uint8_t *skip_m, *skip_o;                      // [2 N_pad], [5 N_pad]: rows whose EQ is frozen (K2 row skip)
    // preset-mute envelope (usb_audio.c:456-498, Q15 use :975-980)
    uint32_t *env;                                 // [5][N_pad] loading, counter, smooth gain (float bits), sample rate, envelope mode on
    int32_t *vol_base, *vol_master;                // [N_pad] host volume Q15 (:975), master_volume_q15
    float *o_glin;                                 // [5][N_pad] outputs[o].gain_linear
    int32_t *pmg;                                  // [N_pad] the constant preset_mute_gain of dspi_chainq_set_params, as Q15 (:976-978)
    int32_t *vmm;                                  // [packets of the call][N_pad] vol_mul_master (:980) of envelope-mode instances
};

// fast_mul_q28(), dsp_pipeline.c:47-58: 32-bit wrapping, lo*lo partial product dropped
__device__ __forceinline__ int32_t mul_q28(int32_t a, int32_t b)
{
    const int32_t ah = a >> 16, bh = b >> 16;
    const uint32_t al = (uint32_t)a & 0xFFFFu, bl = (uint32_t)b & 0xFFFFu;
    const uint32_t high = (uint32_t)ah * (uint32_t)bh;
    const uint32_t mid = (uint32_t)ah * bl + al * (uint32_t)bh;
    return (int32_t)((high << 4) + (uint32_t)((int32_t)mid >> 12));
}
// fast_mul_q15(), config.h:556-567
__device__ __forceinline__ int32_t mul_q15(int32_t s, int32_t g)
{
    const int32_t sh = s >> 16, gh = g >> 16;
    const uint32_t sl = (uint32_t)s & 0xFFFFu, gl = (uint32_t)g & 0xFFFFu;
    const uint32_t hh = (uint32_t)sh * (uint32_t)gh;
    const uint32_t mid = (uint32_t)sh * gl + sl * (uint32_t)gh;
    const uint32_t ll = sl * gl;
    return (int32_t)((hh << 17) + (mid << 1) + (ll >> 15));
}

// leveller.c:124-139 (plain float, the RP2040's soft-float never fuses)
__device__ __forceinline__ float gain_computer(float x_db, float threshold, float ratio, float knee)
{
    const float half_knee = __fmul_rn(knee, 0.5f);
    if (x_db > __fadd_rn(threshold, half_knee)) return 0.0f;
    if (x_db >= __fadd_rn(threshold, -half_knee)) {
        const float dd = __fadd_rn(__fadd_rn(threshold, half_knee), -x_db);
        const float k = __fadd_rn(1.0f, -__fdiv_rn(1.0f, ratio));
        return __fdiv_rn(__fmul_rn(__fmul_rn(k, dd), dd), __fmul_rn(2.0f, knee));
    }
    return __fmul_rn(__fadd_rn(threshold, -x_db), __fadd_rn(1.0f, -__fdiv_rn(1.0f, ratio)));
}

// ---------------------------------------------------------------------------------------------
// pre: PCM unpack + preamp (usb_audio.c:997-1015), loudness TDF2 shelves (:1018-1047) -> master rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
chainq_pre_kernel(ChainQ d, const uint8_t *__restrict__ pcm, uint32_t bit_depth, uint32_t f_begin, uint32_t f_end, uint32_t F)
{
    // warp = 16 instances x {L, R}, lane l = side l >> 4 of instance inst16 + (l & 15): see chain_pre_kernel (chain_f32.cu)
    __shared__ int32_t tile_s[2][32][kXs];                // per warp: [frame][row = side * 16 + instance]
    __shared__ uint32_t pcm_s[2][2][16][49];              // per warp, double-buffered: 16 instances x 32 frames x <= 6 bytes (rows padded to 49 words)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t inst16 = (blockIdx.x * 2 + warp) * 16;
    if (inst16 >= d.N_pad) return;
    const uint32_t side = lane >> 4, li = lane & 15;
    const uint32_t inst = inst16 + li;
    const bool live = inst < d.N;
    const size_t Np = d.N_pad;
    int32_t (*tile)[kXs] = tile_s[warp];

    const bool loud_on = d.flags[inst] & F_LOUD;
    const uint8_t loud_byp = d.loud_byp[inst];
    int32_t lc[2][5], ls[2][2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int k = 0; k < 5; k++) lc[j][k] = d.loud_c[(j * 5 + k) * Np + inst];
        ls[j][0] = d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst];
        ls[j][1] = d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst];
    }
    const int32_t preamp = d.preamp[side * Np + inst];
    const uint32_t bpf = bit_depth == 24 ? 6u : 4u;
    const bool words_ok = ((reinterpret_cast<uintptr_t>(pcm) | ((size_t)F * bpf) | ((size_t)f_begin * bpf)) & 3u) == 0;
    const uint8_t *my_pcm = pcm + (size_t)inst * F * bpf;
    const uint32_t n_inst = min(16u, d.N > inst16 ? d.N - inst16 : 0u);
    auto fetch = [&](uint32_t f0, int buf) {               // see chain_pre_kernel (chain_f32.cu)
        if (words_ok && f0 < f_end) {
            const uint32_t nwords = (min(32u, f_end - f0) * bpf + 3) / 4;
            for (uint32_t i = 0; i < n_inst; i++) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(pcm + ((size_t)(inst16 + i) * F + f0) * bpf);
                for (uint32_t w = lane; w < nwords; w += 32) cp_async_4(&pcm_s[warp][buf][i][w], src + w);
            }
        }
        cp_async_commit();
    };
    fetch(f_begin, 0);

    int buf = 0;
    for (uint32_t f0 = f_begin; f0 < f_end; f0 += 32, buf ^= 1) {
        const uint32_t nv = min(32u, f_end - f0);
        const uint8_t *tile_bytes = my_pcm + (size_t)f0 * bpf;
        fetch(f0 + 32, buf ^ 1);
        if (words_ok) {
            cp_async_wait<1>();
            __syncwarp();
            tile_bytes = reinterpret_cast<const uint8_t *>(pcm_s[warp][buf][li]);
        }
        for (uint32_t t = 0; t < nv; t++) {
            const uint8_t *q = tile_bytes + (size_t)t * bpf;
            int32_t raw = 0;
            if (live) {
                if (bit_depth == 24) {
                    const uint8_t *b = q + side * 3;
                    raw = ((int32_t)((uint32_t)b[2] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[0] << 8)) >> 2;       // :1001
                } else {
                    const uint8_t *b = q + side * 2;
                    raw = (int32_t)((uint32_t)(int32_t)(int16_t)((uint16_t)b[0] | (uint16_t)b[1] << 8) << 14);       // :1010
                }
            }
            int32_t x = mul_q28(raw, preamp);
            if (loud_on) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if ((loud_byp >> j) & 1) continue;
                    const int32_t result = mul_q28(lc[j][0], x) + ls[j][0];                                          // :1026
                    ls[j][0] = mul_q28(lc[j][1], x) - mul_q28(lc[j][3], result) + ls[j][1];
                    ls[j][1] = mul_q28(lc[j][2], x) - mul_q28(lc[j][4], result);
                    x = result;
                }
            }
            tile[t][lane] = x;
        }
        __syncwarp();
        if ((uint32_t)lane < nv) {
#pragma unroll 8
            for (int r = 0; r < 32; r++)
                d.mrow[((size_t)(r >> 4) * Np + inst16 + (r & 15)) * d.ldF + f0 + lane] = tile[lane][r];
        }
        __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
        d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst] = ls[j][0];
        d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst] = ls[j][1];
    }
}

// ---------------------------------------------------------------------------------------------
// post: Q28 leveller (leveller.c:275-389), input peaks, crossfeed (crossfeed.c:161-180), per packet
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
chainq_post_kernel(ChainQ d, uint32_t p0, uint32_t n_packets, uint32_t fpp)
{
    extern __shared__ int32_t smem_q[];                    // per warp: packet columns [fpp][33] + look-ahead reads [fpp][33]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t side = lane >> 4;
    const uint32_t inst16 = (blockIdx.x * (blockDim.x >> 5) + warp) * 16;
    if (inst16 >= d.N_pad) return;
    const uint32_t inst = inst16 + (lane & 15);
    const size_t Np = d.N_pad;
    int32_t *xw = smem_q + (size_t)warp * 2 * fpp * kXs;
    int32_t *xs = xw + lane;
    int32_t *hs = xw + (size_t)fpp * kXs + lane;

    const uint8_t flags = d.flags[inst];
    const bool lev_on = flags & F_LEV, xf_on = flags & F_XFEED, lookahead = flags & F_LOOKAHEAD;
    const int32_t xf_a0 = d.xf[0 * Np + inst], xf_b1 = d.xf[1 * Np + inst], xf_ap = d.xf[4 * Np + inst];
    int32_t xf_lp = d.xf[(2 + side) * Np + inst], xf_as = d.xf[(5 + side) * Np + inst];
    float lvc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) lvc[k] = d.lev_c[k * Np + inst];
    int32_t env = d.lev_i[side * Np + inst];
    int32_t gain_q = d.lev_i[2 * Np + inst], gain_prev_q = d.lev_i[3 * Np + inst];
    float smooth_db = d.lev_f[inst];
    uint32_t la_idx = d.lev_idx[inst];
    int32_t *la_buf = d.lev_la + (size_t)side * kLa * Np + inst;

    int32_t peak_last = 0;
    uint16_t clip = 0;
    for (uint32_t p = p0; p < p0 + n_packets; p++) {
        const uint32_t f0 = p * fpp;
        if (lev_on && lookahead) {                                            // see chain_post_kernel (chain_f32.cu)
            uint32_t idx = la_idx;
            for (uint32_t i = 0; i < fpp; i++) {
                cp_async_4(hs + i * kXs, la_buf + (size_t)idx * Np);
                if (++idx >= (uint32_t)kLa) idx = 0;
            }
        }
        for (int r = 0; r < 32; r++) {
            const int32_t *row = d.mrow + ((size_t)(r >> 4) * Np + inst16 + (r & 15)) * d.ldF + f0;
            for (uint32_t t = lane; t < fpp; t += 32) cp_async_4(xw + t * kXs + r, row + t);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();

        // PASS 3 (:1065-1073) for one sample: input peak, crossfeed (every lane walks the same shuffle)
        int32_t pk = 0;
        auto peak_and_crossfeed = [&](uint32_t i, int32_t v) {
            const int32_t a = abs(v);
            if (a > pk) pk = a;
            int32_t lp = 0, ap = 0;
            if (xf_on) {
                lp = mul_q28(xf_a0, v) + mul_q28(xf_b1, xf_lp);                                           // crossfeed.c:166-167
                xf_lp = lp;
                ap = mul_q28(xf_ap, lp) + xf_as;                                                          // :172
                xf_as = lp - mul_q28(xf_ap, ap);                                                          // :173
            }
            const int32_t ap_other = __shfl_xor_sync(0xffffffffu, ap, 16);
            if (xf_on) v = (v - lp) + ap_other;                                                           // :178-179
            xs[i * kXs] = v;
        };

        // PASS 2.5: leveller (leveller.c:275-389); every lane walks the same shuffles
        if (__any_sync(0xffffffffu, lev_on)) {
            const int32_t a_rms = __float2int_rz(__fmul_rn(lvc[0], 268435456.0f));                       // :286
            const int32_t one_minus = kUnity - a_rms;
            int32_t e = env;
            for (uint32_t i = 0; i < fpp; i++) {                                                         // :292-299
                const int32_t s = xs[i * kXs];
                const int32_t sq = mul_q28(s, s);
                e = mul_q28(a_rms, e) + mul_q28(one_minus, sq);
            }
            const int32_t e_other = __shfl_xor_sync(0xffffffffu, e, 16);
            const float inv_q28 = 1.0f / 268435456.0f;
            const float el = __fmul_rn((float)(side ? e_other : e), inv_q28), er = __fmul_rn((float)(side ? e : e_other), inv_q28);
            const float rms_sq = (el > er) ? el : er;
            const float rms_db = __fmul_rn(10.0f, (float)log10((double)__fadd_rn(rms_sq, 1e-30f)));     // :311 (libm policy)
            float gc_db;
            if (rms_db < lvc[7]) gc_db = 0.0f;
            else {
                gc_db = gain_computer(rms_db, lvc[3], lvc[4], lvc[5]);
                gc_db = __fadd_rn(gc_db, lvc[6]);
                if (gc_db > lvc[8]) gc_db = lvc[8];
            }
            const float alpha_s = (gc_db < smooth_db) ? lvc[1] : lvc[2];
            const float alpha = (float)pow((double)alpha_s, (double)(float)fpp);                          // :327
            const float new_smooth = __fadd_rn(__fmul_rn(alpha, smooth_db), __fmul_rn(__fadd_rn(1.0f, -alpha), gc_db));   // :328-329
            const float gl = (float)pow(10.0, (double)__fdiv_rn(new_smooth, 20.0f));                      // :332
            const int32_t g_cur = __float2int_rz(__fmul_rn(gl, 268435456.0f));                            // :334 (saturating)
            const int32_t g_prev = gain_q;
            // :352 is gain = prev + (int32)((int64)(cur - prev) * i / (fpp - 1)) per sample (C division: towards zero).  With
            // diff = Q * D + R (D = fpp - 1, R with the sign of diff, |R| < D) the quotient is Q * i + trunc(R * i / D), both terms
            // of one sign, so it is carried incrementally: off += Q, acc += R, one correction when |acc| reaches D.  All in
            // 32-bit wrapping arithmetic, which is the int32 cast of :352; no 64-bit division in the per-sample loop.
            const int32_t g_diff = (int32_t)((uint32_t)g_cur - (uint32_t)g_prev);
            const int32_t den = fpp > 1 ? (int32_t)(fpp - 1) : 1;
            const int32_t ramp_q = g_diff / den, ramp_r = g_diff - ramp_q * den;
            uint32_t off = 0;
            int32_t acc = 0;
            // The leveller's per-sample part and PASS 3 share one loop: the ramp, the look-ahead exchange and the peak limit of
            // sample i+1 do not depend on the crossfeed recurrence of sample i, so the two serial chains overlap.
            for (uint32_t i = 0; i < fpp; i++) {                                                          // :347-386, :1065-1073
                int32_t gain = fpp == 1 ? g_cur : (int32_t)((uint32_t)g_prev + off);                      // :352
                off += (uint32_t)ramp_q;
                acc += ramp_r;
                if (acc >= den) { acc -= den; off++; }
                else if (acc <= -den) { acc += den; off--; }
                const int32_t x0 = xs[i * kXs];
                int32_t o = x0;
                if (lev_on && lookahead) {
                    const int32_t held = hs[i * kXs];
                    la_buf[(size_t)la_idx * Np] = o;
                    o = held;
                    la_idx++;
                    if (la_idx >= (uint32_t)kLa) la_idx = 0;
                }
                const int32_t o_other = __shfl_xor_sync(0xffffffffu, o, 16);
                if (gain > kUnity) {                                                                      // :370-379
                    const int32_t ol = side ? o_other : o, orr = side ? o : o_other;
                    float peak = fabsf(__fmul_rn((float)ol, inv_q28));
                    const float pr = fabsf(__fmul_rn((float)orr, inv_q28));
                    if (pr > peak) peak = pr;
                    if (peak > 0.0f) {
                        const float max_g_f = __fdiv_rn(0.70795f, peak);
                        const int32_t max_g = __float2int_rz(__fmul_rn(max_g_f, 268435456.0f));
                        if (max_g < gain) gain = (max_g > kUnity) ? max_g : kUnity;
                    }
                }
                peak_and_crossfeed(i, lev_on ? mul_q28(o, gain) : x0);
            }
            if (lev_on) {
                env = e;
                smooth_db = new_smooth;
                gain_prev_q = g_prev;
                gain_q = g_cur;
            }
        } else {
            for (uint32_t i = 0; i < fpp; i++) peak_and_crossfeed(i, xs[i * kXs]);
        }
        peak_last = pk;
        if (pk > kClipThresh) clip |= (uint16_t)(1u << side);
        __syncwarp();
        for (uint32_t t = lane; t < fpp; t += 32) {
#pragma unroll 8
            for (int r = 0; r < 32; r++)
                d.mrow[((size_t)(r >> 4) * Np + inst16 + (r & 15)) * d.ldF + f0 + t] = xw[t * kXs + r];
        }
        __syncwarp();
    }

    d.xf[(2 + side) * Np + inst] = xf_lp;
    d.xf[(5 + side) * Np + inst] = xf_as;
    d.lev_i[side * Np + inst] = env;
    if (side == 0) {
        d.lev_i[2 * Np + inst] = gain_q;
        d.lev_i[3 * Np + inst] = gain_prev_q;
        d.lev_f[inst] = smooth_db;
        d.lev_idx[inst] = la_idx;
    }
    d.peaks[side * Np + inst] = (uint16_t)(peak_last >> 13);                                              // :1279-1280
    const uint16_t clip_other = (uint16_t)__shfl_xor_sync(0xffffffffu, (uint32_t)clip, 16);
    if (side == 0 && (clip | clip_other)) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), (unsigned int)(clip | clip_other) << (16 * (inst & 1)));
}

// ---------------------------------------------------------------------------------------------
// matrix mix in Q15 (usb_audio.c:1076-1100): lane = frame
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chainq_mix_kernel(ChainQ d, uint32_t f_begin, uint32_t f_end)
{
    const int lane = threadIdx.x & 31;
    constexpr int kB = 4;
    const uint32_t n_tiles = (f_end - f_begin + 32 * kB - 1) / (32 * kB);
    const uint64_t units = (uint64_t)d.N * n_tiles;
    const size_t Np = d.N_pad;
    for (uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t inst = (uint32_t)(u / n_tiles), tile = (uint32_t)(u % n_tiles);
        const uint32_t fbase = f_begin + tile * 32 * kB + lane;
        int32_t l[kB], r[kB];
#pragma unroll
        for (int j = 0; j < kB; j++) {
            const uint32_t f = fbase + 32 * j;
            l[j] = f < f_end ? d.mrow[(size_t)inst * d.ldF + f] : 0;
            r[j] = f < f_end ? d.mrow[(Np + inst) * d.ldF + f] : 0;
        }
#pragma unroll
        for (int o = 0; o < kOuts; o++) {
            const bool enabled = d.o_flags[o * Np + inst] & O_ENABLED;
            const int32_t gl = d.o_gl[o * Np + inst], gr = d.o_gr[o * Np + inst];
#pragma unroll
            for (int j = 0; j < kB; j++) {
                int32_t v = 0;
                if (enabled) {
                    if (gl != 0 && gr != 0) v = mul_q15(l[j], gl) + mul_q15(r[j], gr);
                    else if (gl != 0) v = mul_q15(l[j], gl);
                    else if (gr != 0) v = mul_q15(r[j], gr);
                }
                const uint32_t f = fbase + 32 * j;
                if (f < f_end) d.orow[((size_t)o * Np + inst) * d.ldF + f] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// outputs after the EQ: gain (:1203-1212), delay (:1216-1230), peaks (:1232-1241), 24-bit words (:1243-1257), sub (:1259-1275)
// ---------------------------------------------------------------------------------------------
// update_preset_mute_envelope() (usb_audio.c:466-498) for every packet of the call, one instance per thread, then the Q15
// volume chain of :976-980: pmg = (int32)(g * 32768 + 0.5) clamped, vmm[p] = mul_q15(mul_q15(vol_base, pmg), master_q15)
__global__ void chainq_env_kernel(ChainQ d, uint32_t n_packets, uint32_t fpp)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t Np = d.N_pad;
    if (inst >= d.N || !d.env[4 * Np + inst]) return;
    uint32_t loading = d.env[0 * Np + inst], counter = d.env[1 * Np + inst];
    float g = __uint_as_float(d.env[2 * Np + inst]);
    const uint32_t fs = d.env[3 * Np + inst];
    unsigned long long ts = ((unsigned long long)fs * 8ull + 999ull) / 1000ull;
    if (ts < 1ull) ts = 1ull;
    if (ts > 0xFFFFFFFFull) ts = 0xFFFFFFFFull;
    float step = __fdiv_rn((float)fpp, (float)(uint32_t)ts);
    if (step > 1.0f) step = 1.0f;
    const int32_t vol_base = d.vol_base[inst], master = d.vol_master[inst];
    for (uint32_t p = 0; p < n_packets; p++) {
        const bool active = loading != 0;
        if (active) {
            if (counter > fpp) counter -= fpp;
            else { counter = 0; loading = 0; }
        }
        const float target = active ? 0.0f : 1.0f;
        if (g < target)      { g = __fadd_rn(g, step);  if (g > target) g = target; }
        else if (g > target) { g = __fadd_rn(g, -step); if (g < target) g = target; }
        int32_t pmg = __float2int_rz(__fadd_rn(__fmul_rn(g, 32768.0f), 0.5f));          // :976-978
        if (pmg < 0) pmg = 0;
        if (pmg > 32768) pmg = 32768;
        d.vmm[(size_t)p * Np + inst] = mul_q15(mul_q15(vol_base, pmg), master);         // :979-980
    }
    d.env[0 * Np + inst] = loading;
    d.env[1 * Np + inst] = counter;
    d.env[2 * Np + inst] = __float_as_uint(g);
}

// Mass reconfiguration of the dynamics stages on the device (SURVEY f-1), RP2040 stores: see chain_f32.cu
__global__ void chainq_dynamics_kernel(ChainQ d, uint32_t inst0, uint32_t n, const dspi_dynamics_config *__restrict__ cfgs, float fs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t inst = inst0 + i;
    const size_t Np = d.N_pad;
    const dspi_dynamics_config cfg = cfgs[i];
    uint8_t flags = d.flags[inst] & (uint8_t)~(F_XFEED | F_LEV | F_LOOKAHEAD | F_LOUD);
    float a0, b1, ap;
    const bool xon = dyn::crossfeed_coeffs(cfg.crossfeed, fs, a0, b1, ap);
    const float scale = 268435456.0f;                                        // crossfeed.c:115-118
    d.xf[0 * Np + inst] = xon ? dyn::f2i_sat(__fmul_rn(a0, scale)) : 0;
    d.xf[1 * Np + inst] = xon ? dyn::f2i_sat(__fmul_rn(b1, scale)) : 0;
    d.xf[4 * Np + inst] = xon ? dyn::f2i_sat(__fmul_rn(ap, scale)) : 0;
    d.xf[2 * Np + inst] = 0; d.xf[3 * Np + inst] = 0; d.xf[5 * Np + inst] = 0; d.xf[6 * Np + inst] = 0;
    if (cfg.crossfeed.enabled) flags |= F_XFEED;
    float lv[9];
    dyn::leveller_coeffs(cfg.leveller, fs, lv);
#pragma unroll
    for (int k = 0; k < 9; k++) d.lev_c[k * Np + inst] = lv[k];
    if (cfg.leveller.enabled) flags |= F_LEV;
    if (cfg.leveller.lookahead) flags |= F_LOOKAHEAD;
    uint32_t row;
    const int16_t vol_mul = dyn::host_volume(cfg.volume_8_8, row);
    float lo_db, hi_db;
    dyn::loudness_row_gains((int)row, cfg.loudness_ref_spl, cfg.loudness_intensity_pct, lo_db, hi_db);
    int32_t c[5];
    bool byp;
    uint8_t lb = 0;
    const float lfs = fs < 1.0f ? 48000.0f : fs;
    dyn::shelf_q28(200.0f, 0.707f, lo_db, false, lfs, c, byp);
    if (byp) lb |= 1;
#pragma unroll
    for (int k = 0; k < 5; k++) d.loud_c[(0 * 5 + k) * Np + inst] = c[k];
    dyn::shelf_q28(6000.0f, 0.707f, hi_db, true, lfs, c, byp);
    if (byp) lb |= 2;
#pragma unroll
    for (int k = 0; k < 5; k++) d.loud_c[(1 * 5 + k) * Np + inst] = c[k];
    d.loud_byp[inst] = lb;
    if (cfg.loudness_enabled) flags |= F_LOUD;
    d.flags[inst] = flags;
    const int32_t vol_base = cfg.host_mute ? 0 : (int32_t)vol_mul;           // usb_audio.c:975
    d.vol_base[inst] = vol_base;
    const int32_t vmm = mul_q15(mul_q15(vol_base, d.pmg[inst]), d.vol_master[inst]);    // :979-980
    for (int o = 0; o < kOuts; o++)
        d.o_gain[o * Np + inst] = (d.o_flags[o * Np + inst] & O_MUTE) ? 0 : __float2int_rz(__fmul_rn(d.o_glin[o * Np + inst], (float)vmm));   // :1204-1205
}

__device__ __forceinline__ int32_t outq_gain(int32_t v, bool enabled, int32_t gain)
{
    if (enabled) v = (gain == 0) ? 0 : mul_q15(v, gain);
    return v;
}

struct OutCfgQ {
    bool enabled, pair_off, delay_on, mute;
    int32_t gain;                      // constant gain of the call (no envelope)
    float glin;                        // outputs[o].gain_linear
    const int32_t *vmm;                // envelope mode: vol_mul_master per packet, stride N_pad; else nullptr
    uint32_t fpp; size_t Np;
    uint32_t dl;                       // delay & (MAX - 1): MAX aliases to 0 (SURVEY a-10)
    const int32_t *row;
    const int32_t *ring;
};

__device__ __forceinline__ OutCfgQ outq_cfg(const ChainQ &d, uint32_t o, uint32_t inst, bool any_delay, uint32_t fpp)
{
    OutCfgQ c;
    const size_t Np = d.N_pad;
    const uint8_t of = d.o_flags[o * Np + inst];
    const int32_t dly = d.o_dly[o * Np + inst];
    c.enabled = of & O_ENABLED;
    c.pair_off = of & O_PAIR_OFF;
    c.mute = of & O_MUTE;
    c.gain = d.o_gain[o * Np + inst];
    c.glin = d.o_glin[o * Np + inst];
    c.vmm = d.env[4 * Np + inst] ? d.vmm + inst : nullptr;
    c.fpp = fpp;
    c.Np = Np;
    c.delay_on = any_delay && dly > 0;
    c.dl = (uint32_t)dly & (kMaxDelay - 1);
    c.row = d.orow + ((size_t)o * Np + inst) * d.ldF;
    c.ring = d.dline + ((size_t)o * Np + inst) * kMaxDelay;
    return c;
}

// output gain in force at frame T of the call (usb_audio.c:1204-1205): constant, or following the envelope packet by packet
__device__ __forceinline__ int32_t gainq_at(const OutCfgQ &c, uint32_t T)
{
    if (!c.vmm) return c.gain;
    return c.mute ? 0 : __float2int_rz(__fmul_rn(c.glin, (float)c.vmm[(size_t)(T / c.fpp) * c.Np]));
}

// frame T of the call emits the post-gain sample of frame T - dl: inside the call from the output rows,
// before it from the ring (see chain_f32.cu)
__device__ __forceinline__ int32_t outq_sample(const OutCfgQ &c, uint32_t T, uint32_t widx0)
{
    if (!c.delay_on) return outq_gain(c.row[T], c.enabled, gainq_at(c, T));
    if (T >= c.dl) return outq_gain(c.row[T - c.dl], c.enabled, gainq_at(c, T - c.dl));
    return c.ring[(widx0 + T - c.dl) & (kMaxDelay - 1)];
}

__device__ __forceinline__ int32_t clip_s24(int32_t w) { return w > 0x7FFFFF ? 0x7FFFFF : (w < -0x800000 ? -0x800000 : w); }   // config.h:547-551

__global__ void __launch_bounds__(256)
chainq_outpost_kernel(ChainQ d, uint32_t p0, uint32_t n_packets, uint32_t fpp, uint32_t F, int32_t *__restrict__ spdif_out)
{
    const int lane = threadIdx.x & 31;
    const uint64_t units = (uint64_t)d.N * n_packets;
    const size_t Np = d.N_pad;
    constexpr int kPairs = (kOuts - 1) / 2;
    for (uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t inst = (uint32_t)(u / n_packets), p = p0 + (uint32_t)(u % n_packets);
        const uint32_t f0 = p * fpp;
        const bool last = p == p0 + n_packets - 1;
        const bool any_delay = d.flags[inst] & F_ANY_DELAY;
        const uint32_t widx0 = d.widx_in[inst];
        unsigned int clip = 0;
        for (int k = 0; k <= kPairs; k++) {                                   // the S/PDIF pairs, then the sub alone
            const bool is_sub = k == kPairs;
            const uint32_t oa = 2 * k, ob = is_sub ? oa : oa + 1;
            const OutCfgQ ca = outq_cfg(d, oa, inst, any_delay, fpp), cb = outq_cfg(d, ob, inst, any_delay, fpp);
            int32_t pka = 0, pkb = 0;
            constexpr int kB = 4;
            for (uint32_t tb = lane; tb < fpp; tb += 32 * kB) {
                int32_t xa[kB], xb[kB];
#pragma unroll
                for (int j = 0; j < kB; j++) {
                    const uint32_t t = tb + 32 * j;
                    xa[j] = t < fpp ? outq_sample(ca, f0 + t, widx0) : 0;
                    xb[j] = (!is_sub && t < fpp) ? outq_sample(cb, f0 + t, widx0) : 0;
                }
#pragma unroll
                for (int j = 0; j < kB; j++) {
                    const uint32_t t = tb + 32 * j, T = f0 + t;
                    if (t >= fpp) break;
                    const int32_t aa = abs(xa[j]), ab = abs(xb[j]);
                    if (aa > pka) pka = aa;
                    if (ab > pkb) pkb = ab;
                    if (is_sub) {
                        if (ca.enabled) d.subq[(size_t)inst * d.ldF + T] = xa[j];                          // :1270 pdm_push_sample(buf_out[pdm_out][i])
                    } else if (spdif_out) {
                        int2 w = make_int2(0, 0);
                        if (!ca.pair_off) { w.x = clip_s24((xa[j] + 32) >> 6); w.y = clip_s24((xb[j] + 32) >> 6); }   // :1254-1255
                        *reinterpret_cast<int2 *>(spdif_out + (((size_t)inst * kPairs + k) * F + T) * 2) = w;
                    }
                }
            }
            // the reference compares signed values (`if (a > pk)`), so INT_MIN from abs(INT_MIN) never wins: plain signed max
            pka = __reduce_max_sync(0xffffffffu, pka);
            pkb = __reduce_max_sync(0xffffffffu, pkb);
            if (lane == 0) {
                if (last) {
                    uint16_t pq = (uint16_t)(pka >> 13);                                                  // :1239 / :1267
                    if (is_sub && !ca.enabled) pq = 0;                                                    // :1273
                    d.peaks[(2 + oa) * Np + inst] = pq;
                    if (!is_sub) d.peaks[(2 + ob) * Np + inst] = (uint16_t)(pkb >> 13);
                }
                if (pka > kClipThresh && (!is_sub || ca.enabled)) clip |= 1u << (2 + oa);
                if (!is_sub && pkb > kClipThresh) clip |= 1u << (2 + ob);
            }
        }
        if (lane == 0 && clip) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), clip << (16 * (inst & 1)));
    }
}

__global__ void __launch_bounds__(256)
chainq_ring_kernel(ChainQ d, uint32_t F, uint32_t fpp)
{
    const int lane = threadIdx.x & 31;
    const uint64_t units = (uint64_t)d.N * kOuts;
    const size_t Np = d.N_pad;
    for (uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t inst = (uint32_t)(u / kOuts), o = (uint32_t)(u % kOuts);
        const bool any_delay = d.flags[inst] & F_ANY_DELAY;
        const uint32_t widx0 = d.widx_in[inst];
        const OutCfgQ c = outq_cfg(d, o, inst, any_delay, fpp);
        if (c.delay_on) {
            int32_t *ring = d.dline + ((size_t)o * Np + inst) * kMaxDelay;
            for (uint32_t T = (F > (uint32_t)kMaxDelay ? F - kMaxDelay : 0u) + lane; T < F; T += 32)
                ring[(widx0 + T) & (kMaxDelay - 1)] = outq_gain(c.row[T], c.enabled, gainq_at(c, T));
        }
        if (o == 0 && lane == 0) d.widx_out[inst] = any_delay ? (widx0 + F) & (kMaxDelay - 1) : widx0;
    }
}

__global__ void __launch_bounds__(128)
chainq_pdm_kernel(ChainQ d, uint32_t f_begin, uint32_t f_end, uint32_t F, uint32_t *__restrict__ pdm_out)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= d.N) return;
    if (!(d.flags[inst] & F_SUB_ON)) return;                                                              // usb_audio.c:1261
    pdm_modulate_frames(d.pdm, d.subq + (size_t)inst * d.ldF, 1, d.N_pad, inst, f_begin, f_end, F, pdm_out);
}

// filters[][] of n instances (instance-major AoS, 32-byte records) <-> the mirrors of the two EQ engines
__global__ void chainq_scatter_kernel(const dspi_biquad_q28 *__restrict__ aos, uint32_t inst0, uint32_t n, uint32_t Np, dspi_biquad_q28 *__restrict__ m_aos,
                                      dspi_biquad_q28 *__restrict__ o_aos, int to_mirrors)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kRoles * DSPI_MAX_BANDS) return;
    const uint32_t b = i % DSPI_MAX_BANDS, role = (i / DSPI_MAX_BANDS) % kRoles, inst = inst0 + i / (DSPI_MAX_BANDS * kRoles);
    dspi_biquad_q28 *chain_q = const_cast<dspi_biquad_q28 *>(aos) + ((size_t)inst * kRoles + role) * DSPI_MAX_BANDS + b;
    dspi_biquad_q28 *eng_q = role < 2 ? m_aos + ((size_t)role * Np + inst) * DSPI_MAX_BANDS + b : o_aos + ((size_t)(role - 2) * Np + inst) * DSPI_MAX_BANDS + b;
    if (to_mirrors) *eng_q = *chain_q;
    else *chain_q = *eng_q;
}

__global__ void chainq_status_kernel(ChainQ d, dspi_status_q28 *__restrict__ out)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= d.N) return;
    dspi_status_q28 s;
    for (int r = 0; r < kRoles; r++) s.peaks[r] = d.peaks[r * d.N_pad + inst];
    s.cpu0_load = 0;
    s.cpu1_load = 0;
    s.clip_flags = d.clip[inst];
    out[inst] = s;
}

int fail(int code, const char *fmt, ...)
{
    size_t cap = 0;
    char *buf = error_buffer(&cap);
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, cap, fmt, ap);
    va_end(ap);
    return code;
}

// host copies of the firmware's integer helpers (per-packet scalars are folded on the host)
int32_t h_mul_q15(int32_t s, int32_t g)
{
    const int32_t sh = s >> 16, gh = g >> 16;
    const uint32_t sl = (uint32_t)s & 0xFFFFu, gl = (uint32_t)g & 0xFFFFu;
    const uint32_t hh = (uint32_t)sh * (uint32_t)gh, mid = (uint32_t)sh * gl + sl * (uint32_t)gh, ll = sl * gl;
    return (int32_t)((hh << 17) + (mid << 1) + (ll >> 15));
}
int32_t h_f2i_sat(float x)
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

}  // namespace
}  // namespace dspi

using dspi::ChainQ;
using dspi::fail;

#define CU_OK(expr)                                                                                         \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) return fail(DSPI_ECUDA, "%s -> %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

struct dspi_chainq {
    dspi_chain_desc desc;
    ChainQ d;
    cudaStream_t stream;                  // the engine stream callers see; stages run on st.* between ev_begin and ev_done
    dspi::ChainStreams st;
    dspi_biquad_q28 *d_aos;
    dspi_eq *eq_m, *eq_o;            // K2 engines over the master rows (2 N_pad channels) and the output rows (5 N_pad)
    std::vector<void *> allocs;
    uint64_t launches;
    void *d_pcm; size_t pcm_bytes;
    int32_t *d_spdif; size_t spdif_bytes;
    uint32_t *d_pdmout; size_t pdmout_bytes;
    dspi_status_q28 *d_status;
    uint32_t env_instances;          // instances in envelope mode
    uint32_t vmm_packets;            // capacity of d.vmm in packets
};

namespace {

template <typename T>
cudaError_t dev_alloc(dspi_chainq *c, T **p, size_t count, bool zero = true)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return e;
    c->allocs.push_back(q);
    *p = (T *)q;
    return zero ? cudaMemsetAsync(q, 0, count * sizeof(T), c->stream) : cudaSuccess;
}

cudaError_t init_states(dspi_chainq *c)
{
    const size_t Np = c->d.N_pad;
    std::vector<int32_t> unity(Np, 1 << 28), seed(Np, 123456789);
    cudaError_t e;
    if ((e = cudaMemsetAsync(c->d.lev_i, 0, 4 * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(c->d.lev_i + 2 * Np, unity.data(), Np * 4, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;   // leveller.c:101-102
    if ((e = cudaMemcpyAsync(c->d.lev_i + 3 * Np, unity.data(), Np * 4, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.lev_f, 0, Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.lev_idx, 0, Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.lev_la, 0, (size_t)2 * dspi::kLa * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.loud_st, 0, 8 * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.dline, 0, (size_t)dspi::kOuts * dspi::kMaxDelay * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.widx_in, 0, Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.widx_out, 0, Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.pdm, 0, 9 * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(c->d.pdm + 7 * Np, seed.data(), Np * 4, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.peaks, 0, (size_t)dspi::kRoles * Np * 2, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.clip, 0, Np * 2, c->stream)) != cudaSuccess) return e;
    return cudaStreamSynchronize(c->stream);
}

}  // namespace

extern "C" {

int dspi_chainq_destroy(dspi_chainq *c)
{
    if (!c) return DSPI_OK;
    cudaSetDevice(c->desc.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->st.destroy();
    if (c->eq_m) dspi_eq_destroy(c->eq_m);
    if (c->eq_o) dspi_eq_destroy(c->eq_o);
    for (void *p : c->allocs) cudaFree(p);
    if (c->d_pcm) cudaFree(c->d_pcm);
    if (c->d_spdif) cudaFree(c->d_spdif);
    if (c->d_pdmout) cudaFree(c->d_pdmout);
    if (c->d.vmm) cudaFree(c->d.vmm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    cudaGetLastError();
    return DSPI_OK;
}

int dspi_chainq_create(dspi_chainq **out, const dspi_chain_desc *desc)
{
    if (!out || !desc) return fail(DSPI_EINVAL, "null argument");
    *out = nullptr;
    if (desc->arith != DSPI_ARITH_Q28) return fail(DSPI_EINVAL, "dspi_chainq engines are Q28 (arith 2)");
    if (desc->n_instances == 0 || desc->max_frames == 0) return fail(DSPI_EINVAL, "n_instances and max_frames must be > 0");
    if (desc->n_bands == 0 || desc->n_bands > DSPI_MAX_BANDS) return fail(DSPI_EINVAL, "n_bands must be 1..%d", DSPI_MAX_BANDS);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(DSPI_ENODEV, "no CUDA device (there is no CPU fallback)"); }
    if (desc->device < 0 || desc->device >= ndev) return fail(DSPI_ENODEV, "device %d out of range", desc->device);
    int major = 0;
    CU_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, desc->device));
    if (major != 10) return fail(DSPI_ENODEV, "device %d is not sm_100", desc->device);
    CU_OK(cudaSetDevice(desc->device));
    dspi_chainq *c = new (std::nothrow) dspi_chainq();
    if (!c) return fail(DSPI_ENOMEM, "host allocation failed");
    c->stream = nullptr;
    c->st = dspi::ChainStreams();
    c->eq_m = c->eq_o = nullptr;
    c->d_aos = nullptr; c->launches = 0; c->d_pcm = nullptr; c->pcm_bytes = 0; c->d_spdif = nullptr; c->spdif_bytes = 0;
    c->d_pdmout = nullptr; c->pdmout_bytes = 0; c->d_status = nullptr;
    c->env_instances = 0; c->vmm_packets = 0;
    c->desc = *desc;
    ChainQ &d = c->d;
    memset(&d, 0, sizeof(d));
    d.N = desc->n_instances;
    d.N_pad = (d.N + 31) / 32 * 32;
    d.nb = desc->n_bands;
    d.max_frames = desc->max_frames;
    d.ldF = (d.max_frames + 3u) & ~3u;
    const size_t Np = d.N_pad;
    {
        dspi_eq_desc ed;
        memset(&ed, 0, sizeof(ed));
        ed.arith = DSPI_ARITH_Q28; ed.n_bands = desc->n_bands; ed.device = desc->device;
        ed.n_channels = 2 * d.N_pad;
        int rc = dspi_eq_create(&c->eq_m, &ed);
        ed.n_channels = dspi::kOuts * d.N_pad;
        if (rc == DSPI_OK) rc = dspi_eq_create(&c->eq_o, &ed);
        if (rc != DSPI_OK) { dspi_chainq_destroy(c); return rc; }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = c->st.create(desc->device, desc->n_instances);
#define TRY(x) if (e == cudaSuccess) e = (x)
    TRY(dev_alloc(c, &c->d_aos, Np * dspi::kRoles * DSPI_MAX_BANDS));
    TRY(dev_alloc(c, &d.preamp, 2 * Np));
    TRY(dev_alloc(c, &d.flags, Np));
    TRY(dev_alloc(c, &d.loud_c, 10 * Np));
    TRY(dev_alloc(c, &d.loud_st, 8 * Np));
    TRY(dev_alloc(c, &d.loud_byp, Np));
    TRY(dev_alloc(c, &d.xf, 7 * Np));
    TRY(dev_alloc(c, &d.lev_c, 9 * Np));
    TRY(dev_alloc(c, &d.lev_i, 4 * Np));
    TRY(dev_alloc(c, &d.lev_f, Np));
    TRY(dev_alloc(c, &d.lev_idx, Np));
    TRY(dev_alloc(c, &d.lev_la, (size_t)2 * dspi::kLa * Np));
    TRY(dev_alloc(c, &d.o_gl, dspi::kOuts * Np));
    TRY(dev_alloc(c, &d.o_gr, dspi::kOuts * Np));
    TRY(dev_alloc(c, &d.o_gain, dspi::kOuts * Np));
    TRY(dev_alloc(c, &d.o_flags, dspi::kOuts * Np));
    TRY(dev_alloc(c, &d.o_dly, dspi::kOuts * Np));
    TRY(dev_alloc(c, &d.dline, (size_t)dspi::kOuts * dspi::kMaxDelay * Np));
    TRY(dev_alloc(c, &d.widx_in, Np));
    TRY(dev_alloc(c, &d.widx_out, Np));
    TRY(dev_alloc(c, &d.pdm, 9 * Np));
    TRY(dev_alloc(c, &d.peaks, dspi::kRoles * Np));
    TRY(dev_alloc(c, &d.clip, Np));
    TRY(dev_alloc(c, &d.mrow, (size_t)2 * Np * d.ldF));
    TRY(dev_alloc(c, &d.orow, (size_t)dspi::kOuts * Np * d.ldF));
    TRY(dev_alloc(c, &d.subq, (size_t)Np * d.ldF));
    TRY(dev_alloc(c, &d.skip_m, 2 * Np));
    TRY(dev_alloc(c, &d.skip_o, dspi::kOuts * Np));
    TRY(dev_alloc(c, &c->d_status, Np));
    TRY(dev_alloc(c, &d.env, 5 * Np));
    TRY(dev_alloc(c, &d.vol_base, Np));
    TRY(dev_alloc(c, &d.vol_master, Np));
    TRY(dev_alloc(c, &d.o_glin, dspi::kOuts * Np));
    TRY(dev_alloc(c, &d.pmg, Np));
    // every band of every channel starts bypassed (dsp_init_default_filters, dsp_pipeline.c:177-199)
    if (e == cudaSuccess) {
        std::vector<dspi_biquad_q28> byp(Np * dspi::kRoles * DSPI_MAX_BANDS);
        memset(byp.data(), 0, byp.size() * sizeof(dspi_biquad_q28));
        for (auto &q : byp) q.bypass = 1;
        e = cudaMemcpyAsync(dspi::eq_aos_mirror(c->eq_m), byp.data(), 2 * Np * DSPI_MAX_BANDS * sizeof(dspi_biquad_q28), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dspi::eq_aos_mirror(c->eq_o), byp.data(), dspi::kOuts * Np * DSPI_MAX_BANDS * sizeof(dspi_biquad_q28), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_aos, byp.data(), byp.size() * sizeof(dspi_biquad_q28), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess && (dspi::eq_pack_range(c->eq_m, 0, 2 * d.N_pad, c->stream) || dspi::eq_pack_range(c->eq_o, 0, dspi::kOuts * d.N_pad, c->stream)))
            e = cudaErrorUnknown;
    }
    TRY(init_states(c));
#undef TRY
    if (e != cudaSuccess) {
        fail(e == cudaErrorMemoryAllocation ? DSPI_ENOMEM : DSPI_ECUDA, "chainq setup: %s", cudaGetErrorString(e));
        dspi_chainq_destroy(c);
        return e == cudaErrorMemoryAllocation ? DSPI_ENOMEM : DSPI_ECUDA;
    }
    *out = c;
    return DSPI_OK;
}

int dspi_chainq_reset_state(dspi_chainq *c)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(init_states(c));
    return DSPI_OK;
}

int dspi_chainq_set_params(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_chain_params_q28 *params)
{
    if (!c || !params) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const ChainQ &d = c->d;
    const size_t Np = d.N_pad;
    const int O = dspi::kOuts;
    std::vector<int32_t> preamp(2 * n), loud_c(10 * n), xf(7 * n), gl(O * n), gr(O * n), gain(O * n), dly(O * n);
    std::vector<float> lev_c(9 * n), glin(O * n);
    std::vector<int32_t> vbase(n), vmaster(n), pmgv(n);
    std::vector<uint8_t> flags(n), loud_byp(n), oflags(O * n), skip_m(2 * n), skip_o(O * n);
    std::vector<int32_t> xf_cur(7 * n);
    CU_OK(cudaMemcpy2DAsync(xf_cur.data(), (size_t)n * 4, d.xf + inst0, Np * 4, (size_t)n * 4, 7, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; i++) {
        const dspi_chain_params_q28 &p = params[i];
        int32_t vol_mul = p.host_mute ? 0 : (int32_t)p.host_vol_mul;                                    // usb_audio.c:975
        vbase[i] = vol_mul;
        vmaster[i] = p.master_volume_q15;
        int32_t pmg = (int32_t)(p.preset_mute_gain * 32768.0f + 0.5f);                                  // :976-978
        if (pmg < 0) pmg = 0;
        if (pmg > 32768) pmg = 32768;
        pmgv[i] = pmg;
        vol_mul = dspi::h_mul_q15(vol_mul, pmg);                                                        // :979
        const int32_t vol_mul_master = dspi::h_mul_q15(vol_mul, p.master_volume_q15);                   // :980
        preamp[0 * n + i] = p.preamp_q28[0];
        preamp[1 * n + i] = p.preamp_q28[1];
        bool any_delay = false;
        for (int o = 0; o < O; o++) {
            const dspi_output_channel &oc = p.matrix.outputs[o];
            const dspi_matrix_crosspoint &xl = p.matrix.crosspoints[0][o], &xr = p.matrix.crosspoints[1][o];
            gl[o * n + i] = xl.enabled ? dspi::h_f2i_sat((xl.phase_invert ? -xl.gain_linear : xl.gain_linear) * 32768.0f) : 0;   // :1084-1085
            gr[o * n + i] = xr.enabled ? dspi::h_f2i_sat((xr.phase_invert ? -xr.gain_linear : xr.gain_linear) * 32768.0f) : 0;
            gain[o * n + i] = oc.mute ? 0 : dspi::h_f2i_sat(oc.gain_linear * (float)vol_mul_master);    // :1204-1205
            glin[o * n + i] = oc.gain_linear;
            uint8_t f = (oc.enabled ? dspi::O_ENABLED : 0) | (oc.mute ? dspi::O_MUTE : 0);
            if (o < O - 1 && !oc.enabled && !p.matrix.outputs[o ^ 1].enabled) f |= dspi::O_PAIR_OFF;    // :1248-1251
            oflags[o * n + i] = f;
            skip_o[o * n + i] = (oc.enabled && !oc.mute && !p.bypass_master_eq) ? 0 : 1;                 // :1197-1201 (quirk: gated on bypass_master_eq too)
            int32_t ds = oc.delay_samples;
            if (ds > DSPI_CHAINQ_MAX_DELAY) ds = DSPI_CHAINQ_MAX_DELAY;
            if (ds < 0) ds = 0;
            dly[o * n + i] = ds;
            if (ds > 0) any_delay = true;
        }
        flags[i] = (p.bypass_master_eq ? dspi::F_BYPASS_MASTER : 0) | (p.loudness_enabled ? dspi::F_LOUD : 0) |
                   (p.crossfeed_enabled ? dspi::F_XFEED : 0) | (p.leveller_enabled ? dspi::F_LEV : 0) |
                   (p.leveller_lookahead ? dspi::F_LOOKAHEAD : 0) | (any_delay ? dspi::F_ANY_DELAY : 0) |
                   (p.matrix.outputs[O - 1].enabled ? dspi::F_SUB_ON : 0);
        skip_m[0 * n + i] = skip_m[1 * n + i] = p.bypass_master_eq ? 1 : 0;                              // :1050-1055
        loud_byp[i] = (p.loudness[0].bypass ? 1 : 0) | (p.loudness[1].bypass ? 2 : 0);
        for (int j = 0; j < 2; j++) {
            const int32_t v[5] = { p.loudness[j].b0, p.loudness[j].b1, p.loudness[j].b2, p.loudness[j].a1, p.loudness[j].a2 };
            for (int k = 0; k < 5; k++) loud_c[(j * 5 + k) * n + i] = v[k];
        }
        const int32_t xv[7] = { p.crossfeed.lp_a0, p.crossfeed.lp_b1, p.crossfeed.lp_state_L, p.crossfeed.lp_state_R,
                                p.crossfeed.ap_a, p.crossfeed.ap_state_L, p.crossfeed.ap_state_R };
        // crossfeed_compute_coefficients() is the only writer of crossfeed_state in the firmware and it clears the filter
        // state (crossfeed.c:35-127); a volume / mute / matrix update never touches it.  So the record's state rows are
        // taken only when its coefficients differ from the ones in force; otherwise the running state is kept.
        const bool xf_same = xv[0] == xf_cur[0 * n + i] && xv[1] == xf_cur[1 * n + i] && xv[4] == xf_cur[4 * n + i];
        for (int k = 0; k < 7; k++) {
            const bool is_state = k == 2 || k == 3 || k == 5 || k == 6;
            xf[k * n + i] = (is_state && xf_same) ? xf_cur[k * n + i] : xv[k];
        }
        const float *lv = &p.leveller.alpha_rms;
        for (int k = 0; k < 9; k++) lev_c[k * n + i] = lv[k];
    }
    auto put = [&](void *dst_base, const void *src, int rows, size_t elem) -> cudaError_t {
        return cudaMemcpy2DAsync((char *)dst_base + (size_t)inst0 * elem, Np * elem, src, (size_t)n * elem, (size_t)n * elem, rows,
                                 cudaMemcpyHostToDevice, c->stream);
    };
    CU_OK(put(d.preamp, preamp.data(), 2, 4));
    CU_OK(put(d.flags, flags.data(), 1, 1));
    CU_OK(put(d.loud_c, loud_c.data(), 10, 4));
    CU_OK(put(d.loud_byp, loud_byp.data(), 1, 1));
    CU_OK(put(d.xf, xf.data(), 7, 4));
    CU_OK(put(d.lev_c, lev_c.data(), 9, 4));
    CU_OK(put(d.o_gl, gl.data(), O, 4));
    CU_OK(put(d.o_gr, gr.data(), O, 4));
    CU_OK(put(d.o_gain, gain.data(), O, 4));
    CU_OK(put(d.o_glin, glin.data(), O, 4));
    CU_OK(put(d.vol_base, vbase.data(), 1, 4));
    CU_OK(put(d.vol_master, vmaster.data(), 1, 4));
    CU_OK(put(d.pmg, pmgv.data(), 1, 4));
    CU_OK(put(d.o_flags, oflags.data(), O, 1));
    CU_OK(put(d.o_dly, dly.data(), O, 4));
    CU_OK(put(d.skip_m, skip_m.data(), 2, 1));
    CU_OK(put(d.skip_o, skip_o.data(), O, 1));
    CU_OK(cudaStreamSynchronize(c->stream));
    int rc = dspi::eq_set_skip(c->eq_m, d.skip_m, c->stream);
    if (rc == DSPI_OK) rc = dspi::eq_set_skip(c->eq_o, d.skip_o, c->stream);
    return rc;
}

/* preset-mute envelope of instances [inst0, inst0+n): states == NULL leaves envelope mode (the constant
 * preset_mute_gain of dspi_chainq_set_params applies again) */
int dspi_chainq_set_preset_mute(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_preset_mute *states, uint32_t sample_rate_hz)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t Np = c->d.N_pad;
    std::vector<uint32_t> cur((size_t)n), rows((size_t)5 * n, 0u);
    CU_OK(cudaMemcpyAsync(cur.data(), c->d.env + 4 * Np + inst0, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; i++) {
        if (cur[i]) c->env_instances--;
        if (states) {
            rows[0 * n + i] = states[i].loading ? 1u : 0u;
            rows[1 * n + i] = states[i].counter;
            memcpy(&rows[2 * n + i], &states[i].smooth_gain, 4);
            rows[3 * n + i] = sample_rate_hz;
            rows[4 * n + i] = 1u;
            c->env_instances++;
        }
    }
    CU_OK(cudaMemcpy2DAsync(c->d.env + inst0, Np * 4, rows.data(), (size_t)n * 4, (size_t)n * 4, 5, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chainq_get_preset_mute(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_preset_mute *states)
{
    if (!c || !states) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t Np = c->d.N_pad;
    std::vector<uint32_t> rows((size_t)3 * n);
    CU_OK(cudaMemcpy2DAsync(rows.data(), (size_t)n * 4, c->d.env + inst0, Np * 4, (size_t)n * 4, 3, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; i++) {
        memset(&states[i], 0, sizeof(states[i]));
        states[i].loading = (uint8_t)rows[0 * n + i];
        states[i].counter = rows[1 * n + i];
        memcpy(&states[i].smooth_gain, &rows[2 * n + i], 4);
    }
    return DSPI_OK;
}

/* crossfeed / leveller / loudness coefficients and the host volume of instances [inst0, inst0+n) generated ON THE GPU */
int dspi_chainq_set_dynamics_device(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_dynamics_config *cfgs, float sample_rate)
{
    if (!c || !cfgs) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    dspi_dynamics_config *d_cfg = nullptr;
    CU_OK(cudaMalloc((void **)&d_cfg, (size_t)n * sizeof(*cfgs)));
    cudaError_t e = cudaMemcpyAsync(d_cfg, cfgs, (size_t)n * sizeof(*cfgs), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        dspi::chainq_dynamics_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->d, inst0, n, d_cfg, sample_rate);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_cfg);
    if (e != cudaSuccess) return fail(DSPI_ECUDA, "dynamics coefficient generation: %s", cudaGetErrorString(e));
    c->launches++;
    return DSPI_OK;
}

int dspi_chainq_upload_biquads(dspi_chainq *c, uint32_t inst0, uint32_t n, const dspi_biquad_q28 *biquads)
{
    if (!c || !biquads) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t row = (size_t)dspi::kRoles * DSPI_MAX_BANDS;
    CU_OK(cudaMemcpyAsync(c->d_aos + inst0 * row, biquads, n * row * sizeof(dspi_biquad_q28), cudaMemcpyHostToDevice, c->stream));
    const uint32_t Np = c->d.N_pad, items = n * dspi::kRoles * DSPI_MAX_BANDS;
    dspi::chainq_scatter_kernel<<<(items + 255) / 256, 256, 0, c->stream>>>(c->d_aos, inst0, n, Np, (dspi_biquad_q28 *)dspi::eq_aos_mirror(c->eq_m),
                                                                           (dspi_biquad_q28 *)dspi::eq_aos_mirror(c->eq_o), 1);
    CU_OK(cudaGetLastError());
    c->launches++;
    for (int role = 0; role < dspi::kRoles; role++) {
        int rc = role < 2 ? dspi::eq_pack_range(c->eq_m, role * Np + inst0, n, c->stream)
                          : dspi::eq_pack_range(c->eq_o, (role - 2) * Np + inst0, n, c->stream);
        if (rc) return rc;
    }
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chainq_set_eq_params_device(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_eq_param *recipes, float sample_rate)
{
    if (!c || !recipes) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    // the sub-engines generate and pack on their own streams: everything issued on the engine stream so far (an
    // asynchronous process_device in particular) must have finished reading the coefficient stores first
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(cudaStreamSynchronize(c->stream));
    const uint32_t Np = c->d.N_pad;
    std::vector<dspi_eq_param> tmp((size_t)n * DSPI_MAX_BANDS);
    for (int role = 0; role < dspi::kRoles; role++) {               // filter_recipes[role][band] of every instance -> one engine range per role
        for (uint32_t i = 0; i < n; i++)
            memcpy(&tmp[(size_t)i * DSPI_MAX_BANDS], &recipes[((size_t)i * dspi::kRoles + role) * DSPI_MAX_BANDS], DSPI_MAX_BANDS * sizeof(dspi_eq_param));
        int rc = role < 2 ? dspi_eq_set_params_device(c->eq_m, role * Np + inst0, n, tmp.data(), sample_rate)
                          : dspi_eq_set_params_device(c->eq_o, (role - 2) * Np + inst0, n, tmp.data(), sample_rate);
        if (rc) return rc;
        for (uint32_t i = 0; i < n; i++)                            // the clamps, written back like the reference does
            memcpy(&recipes[((size_t)i * dspi::kRoles + role) * DSPI_MAX_BANDS], &tmp[(size_t)i * DSPI_MAX_BANDS], DSPI_MAX_BANDS * sizeof(dspi_eq_param));
    }
    return DSPI_OK;
}

int dspi_chainq_download_biquads(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_biquad_q28 *biquads)
{
    if (!c || !biquads) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t row = (size_t)dspi::kRoles * DSPI_MAX_BANDS;
    const uint32_t Np = c->d.N_pad, items = n * dspi::kRoles * DSPI_MAX_BANDS;
    for (int role = 0; role < dspi::kRoles; role++) {
        int rc = role < 2 ? dspi::eq_unpack_range(c->eq_m, role * Np + inst0, n, c->stream)
                          : dspi::eq_unpack_range(c->eq_o, (role - 2) * Np + inst0, n, c->stream);
        if (rc) return rc;
    }
    dspi::chainq_scatter_kernel<<<(items + 255) / 256, 256, 0, c->stream>>>(c->d_aos, inst0, n, Np, (dspi_biquad_q28 *)dspi::eq_aos_mirror(c->eq_m),
                                                                           (dspi_biquad_q28 *)dspi::eq_aos_mirror(c->eq_o), 0);
    CU_OK(cudaGetLastError());
    c->launches++;
    CU_OK(cudaMemcpyAsync(biquads, c->d_aos + inst0 * row, n * row * sizeof(dspi_biquad_q28), cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chainq_process_device(dspi_chainq *c, const void *d_pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t fpp, int32_t *d_spdif,
                               uint32_t *d_pdm, dspi_status_q28 *d_status)
{
    if (!c || !d_pcm) return fail(DSPI_EINVAL, "null argument");
    if (bit_depth != 16 && bit_depth != 24) return fail(DSPI_EINVAL, "bit_depth must be 16 or 24");
    if (fpp == 0 || fpp > DSPI_PACKET_MAX) return fail(DSPI_EINVAL, "frames_per_packet must be 1..%d", DSPI_PACKET_MAX);
    if (n_packets == 0) return fail(DSPI_EINVAL, "n_packets must be > 0");
    if ((uint64_t)n_packets * fpp > c->desc.max_frames) return fail(DSPI_ERANGE, "%u frames exceed max_frames %u", n_packets * fpp, c->desc.max_frames);
    CU_OK(cudaSetDevice(c->desc.device));
    const uint32_t F = n_packets * fpp;
    const size_t post_smem = (size_t)4 * 2 * fpp * dspi::kXs * 4;           // 4 warps x (packet + look-ahead columns)
    static dspi::PerDeviceOnce once;
    int dev = 0;
    if (once.needs(&dev)) {
        CU_OK(cudaFuncSetAttribute(dspi::chainq_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)4 * 2 * dspi::kPkt * dspi::kXs * 4)));
        once.mark(dev);
    }
    // Stage pipeline over packet slices on three streams (chain_streams.cuh), stages as in chain_f32.cu.
    dspi::ChainStreams &st = c->st;
    uint32_t slice_bounds[dspi::ChainStreams::kMaxSlices + 1];
    const uint32_t n_slices = (uint32_t)dspi::ChainStreams::plan_slices(n_packets, slice_bounds);
    if (c->env_instances) {                                                  // preset-mute envelope: this call's per-packet volumes
        if (c->vmm_packets < n_packets) {
            CU_OK(cudaStreamSynchronize(c->stream));
            if (c->d.vmm) CU_OK(cudaFree(c->d.vmm));
            c->d.vmm = nullptr; c->vmm_packets = 0;
            CU_OK(cudaMalloc((void **)&c->d.vmm, (size_t)n_packets * c->d.N_pad * sizeof(int32_t)));
            c->vmm_packets = n_packets;
        }
        dspi::chainq_env_kernel<<<(c->d.N + 127) / 128, 128, 0, c->stream>>>(c->d, n_packets, fpp);
        CU_OK(cudaGetLastError());
        c->launches++;
    }
    const ChainQ d = c->d;
    const uint32_t n_sms = st.rest_sms ? st.rest_sms : 148;     // SMs the streaming stages run on (chain_streams.cuh)
    static const uint32_t kStreamCtas = [] { const char *e = getenv("DSPI_CHAIN_CTAS"); const int v = e ? atoi(e) : 0; return (uint32_t)(v >= 1 && v <= 8 ? v : 8); }();   // streaming CTAs (256 threads) per SM
    CU_OK(cudaEventRecord(st.ev_begin, c->stream));
    CU_OK(cudaStreamWaitEvent(st.s_front, st.ev_begin, 0));
    for (uint32_t sl = 0; sl < n_slices; sl++) {
        const uint32_t p0 = slice_bounds[sl], p1 = slice_bounds[sl + 1];
        const uint32_t fb = p0 * fpp, fe = p1 * fpp;
        int rc;
        dspi::chainq_pre_kernel<<<(d.N_pad / 16 + 1) / 2, 64, 0, st.s_front>>>(d, (const uint8_t *)d_pcm, bit_depth, fb, fe, F);
        CU_OK(cudaGetLastError());
        if ((rc = dspi::eq_process_on(c->eq_m, d.mrow + fb, fe - fb, d.ldF, st.s_front)) != DSPI_OK) return rc;
        dspi::chainq_post_kernel<<<(d.N_pad / 16 + 3) / 4, 128, post_smem, st.s_front>>>(d, p0, p1 - p0, fpp);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(st.ev_front[sl], st.s_front));
        CU_OK(cudaStreamWaitEvent(st.s_out, st.ev_front[sl], 0));
        dspi::chainq_mix_kernel<<<n_sms * kStreamCtas, 256, 0, st.s_out>>>(d, fb, fe);
        CU_OK(cudaGetLastError());
        if ((rc = dspi::eq_process_on(c->eq_o, d.orow + fb, fe - fb, d.ldF, st.s_out)) != DSPI_OK) return rc;
        dspi::chainq_outpost_kernel<<<n_sms * kStreamCtas, 256, 0, st.s_out>>>(d, p0, p1 - p0, fpp, F, d_spdif);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(st.ev_out[sl], st.s_out));
        CU_OK(cudaStreamWaitEvent(st.s_pdm, st.ev_out[sl], 0));
        dspi::chainq_pdm_kernel<<<(d.N + 127) / 128, 128, 0, st.s_pdm>>>(d, fb, fe, F, d_pdm);
        CU_OK(cudaGetLastError());
        c->launches += 5;
    }
    dspi::chainq_ring_kernel<<<n_sms * kStreamCtas, 256, 0, st.s_out>>>(d, F, fpp);            // after the last outpost launch (stream order)
    CU_OK(cudaGetLastError());
    c->launches++;
    std::swap(c->d.widx_in, c->d.widx_out);
    CU_OK(cudaEventRecord(st.ev_aux, st.s_out));
    CU_OK(cudaStreamWaitEvent(c->stream, st.ev_aux, 0));
    CU_OK(cudaEventRecord(st.ev_done, st.s_pdm));
    CU_OK(cudaStreamWaitEvent(c->stream, st.ev_done, 0));
    if (d_status) {
        dspi::chainq_status_kernel<<<(c->d.N + 127) / 128, 128, 0, c->stream>>>(c->d, d_status);
        CU_OK(cudaGetLastError());
        c->launches++;
    }
    return DSPI_OK;
}

int dspi_chainq_process_host(dspi_chainq *c, const void *pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t fpp, int32_t *spdif_out,
                             uint32_t *pdm_out, dspi_status_q28 *status)
{
    if (!c || !pcm) return fail(DSPI_EINVAL, "null argument");
    if (bit_depth != 16 && bit_depth != 24) return fail(DSPI_EINVAL, "bit_depth must be 16 or 24");
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t N = c->desc.n_instances, F = (size_t)n_packets * fpp;
    const size_t in_bytes = N * F * (bit_depth == 24 ? 6 : 4), sp_bytes = N * 2 * F * 2 * 4, pd_bytes = N * F * 8 * 4;
    if (in_bytes > c->pcm_bytes) { if (c->d_pcm) cudaFree(c->d_pcm); c->d_pcm = nullptr; c->pcm_bytes = 0; CU_OK(cudaMalloc(&c->d_pcm, in_bytes)); c->pcm_bytes = in_bytes; }
    if (spdif_out && sp_bytes > c->spdif_bytes) { if (c->d_spdif) cudaFree(c->d_spdif); c->d_spdif = nullptr; c->spdif_bytes = 0; CU_OK(cudaMalloc((void **)&c->d_spdif, sp_bytes)); c->spdif_bytes = sp_bytes; }
    if (pdm_out && pd_bytes > c->pdmout_bytes) { if (c->d_pdmout) cudaFree(c->d_pdmout); c->d_pdmout = nullptr; c->pdmout_bytes = 0; CU_OK(cudaMalloc((void **)&c->d_pdmout, pd_bytes)); c->pdmout_bytes = pd_bytes; CU_OK(cudaMemsetAsync(c->d_pdmout, 0, pd_bytes, c->stream)); }
    CU_OK(cudaMemcpyAsync(c->d_pcm, pcm, in_bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = dspi_chainq_process_device(c, c->d_pcm, bit_depth, n_packets, fpp, spdif_out ? c->d_spdif : nullptr, pdm_out ? c->d_pdmout : nullptr,
                                        status ? c->d_status : nullptr);
    if (rc) return rc;
    if (spdif_out) CU_OK(cudaMemcpyAsync(spdif_out, c->d_spdif, sp_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (pdm_out) CU_OK(cudaMemcpyAsync(pdm_out, c->d_pdmout, pd_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (status) CU_OK(cudaMemcpyAsync(status, c->d_status, N * sizeof(dspi_status_q28), cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}


// ---- checkpoint / resume: everything a later process call depends on besides the parameters ----------------
static void state_sections(dspi_chainq *c, std::vector<std::pair<void *, size_t>> &v)
{
    const size_t Np = c->d.N_pad;
    v.push_back({ c->d.loud_st, 8 * Np * 4 });
    v.push_back({ c->d.xf, 7 * Np * 4 });
    v.push_back({ c->d.lev_i, 4 * Np * 4 });
    v.push_back({ c->d.lev_f, Np * 4 });
    v.push_back({ c->d.lev_idx, Np * 4 });
    v.push_back({ c->d.lev_la, (size_t)2 * dspi::kLa * Np * 4 });
    v.push_back({ c->d.dline, (size_t)dspi::kOuts * dspi::kMaxDelay * Np * 4 });
    v.push_back({ c->d.widx_in, Np * 4 });
    v.push_back({ c->d.pdm, 9 * Np * 4 });
    v.push_back({ c->d.peaks, (size_t)dspi::kRoles * Np * 2 });
    v.push_back({ c->d.clip, Np * 2 });
    v.push_back({ c->d.env, 5 * Np * 4 });                                 // preset-mute envelope state and mode
    dspi::eq_state_sections(c->eq_m, v);
    dspi::eq_state_sections(c->eq_o, v);
}

struct StateHeader { uint32_t magic, version, arith, n_instances, n_bands, n_sections; uint64_t bytes; };
static const uint32_t kStateMagic = 0x53505344u;          // "DSPS"

size_t dspi_chainq_state_size(dspi_chainq *c)
{
    if (!c) return 0;
    std::vector<std::pair<void *, size_t>> v;
    state_sections(c, v);
    size_t n = sizeof(StateHeader);
    for (auto &s : v) n += s.second;
    return n;
}

int dspi_chainq_state_export(dspi_chainq *c, void *blob, size_t cap)
{
    if (!c || !blob) return fail(DSPI_EINVAL, "null argument");
    const size_t need = dspi_chainq_state_size(c);
    if (cap < need) return fail(DSPI_ERANGE, "state blob needs %zu bytes, %zu given", need, cap);
    CU_OK(cudaSetDevice(c->desc.device));
    std::vector<std::pair<void *, size_t>> v;
    state_sections(c, v);
    StateHeader h = { kStateMagic, 1u, c->desc.arith, c->desc.n_instances, c->desc.n_bands, (uint32_t)v.size(), (uint64_t)need };
    memcpy(blob, &h, sizeof(h));
    char *p = (char *)blob + sizeof(h);
    for (auto &s : v) {
        CU_OK(cudaMemcpyAsync(p, s.first, s.second, cudaMemcpyDeviceToHost, c->stream));
        p += s.second;
    }
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chainq_state_import(dspi_chainq *c, const void *blob, size_t len)
{
    if (!c || !blob) return fail(DSPI_EINVAL, "null argument");
    std::vector<std::pair<void *, size_t>> v;
    state_sections(c, v);
    StateHeader h;
    if (len < sizeof(h)) return fail(DSPI_EINVAL, "state blob too short");
    memcpy(&h, blob, sizeof(h));
    if (h.magic != kStateMagic || h.version != 1u) return fail(DSPI_EINVAL, "not a dspi_b200 state blob (magic %08x version %u)", h.magic, h.version);
    if (h.arith != c->desc.arith || h.n_instances != c->desc.n_instances || h.n_bands != c->desc.n_bands || h.n_sections != v.size() ||
        h.bytes != dspi_chainq_state_size(c) || len < h.bytes)
        return fail(DSPI_EINVAL, "state blob belongs to a different engine shape (%u instances, arith %u, %llu bytes)", h.n_instances, h.arith,
                    (unsigned long long)h.bytes);
    CU_OK(cudaSetDevice(c->desc.device));
    const char *p = (const char *)blob + sizeof(h);
    for (auto &s : v) {
        CU_OK(cudaMemcpyAsync(s.first, p, s.second, cudaMemcpyHostToDevice, c->stream));
        p += s.second;
    }
    CU_OK(cudaStreamSynchronize(c->stream));
    int rc = dspi::eq_state_imported(c->eq_m, c->stream);
    if (rc == DSPI_OK) rc = dspi::eq_state_imported(c->eq_o, c->stream);
    if (rc) return rc;
    std::vector<uint32_t> on(c->d.N);
    CU_OK(cudaMemcpyAsync(on.data(), c->d.env + (size_t)4 * c->d.N_pad, (size_t)c->d.N * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    c->env_instances = 0;
    for (uint32_t v : on) c->env_instances += v ? 1u : 0u;
    return DSPI_OK;
}

int dspi_chainq_sync(dspi_chainq *c)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

/* SMs reserved for the modulator / left to every other stage (0, 0: no partition, see chain_streams.cuh) */
int dspi_chainq_sm_partition(dspi_chainq *c, uint32_t *pdm_sms, uint32_t *rest_sms)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    if (pdm_sms) *pdm_sms = c->st.pdm_sms;
    if (rest_sms) *rest_sms = c->st.rest_sms;
    return DSPI_OK;
}

void *dspi_chainq_stream(dspi_chainq *c) { return c ? (void *)c->stream : nullptr; }
uint64_t dspi_chainq_launch_count(dspi_chainq *c)
{
    return c ? c->launches + dspi_eq_launch_count(c->eq_m) + dspi_eq_launch_count(c->eq_o) : 0;
}

}  // extern "C"
