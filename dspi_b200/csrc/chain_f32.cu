// chain_f32.cu — K3..K9: the whole per-packet float signal chain of one DSPi device, for thousands
// of independent device instances, sm_100a.
//
// Reference: process_audio_packet(), firmware/DSPi/usb_audio.c:500-1317 — float pipeline :560-967,
// single-core branch :874-960; crossfeed.c:132-156; leveller.c:148-262; pdm_generator.c:351-397.
// Chain order (the code's, not the README's): preamp -> loudness -> master EQ -> leveller ->
// crossfeed (+ input peaks) -> matrix -> per-output EQ -> gain x volume -> delay -> peaks ->
// 24-bit words / delta-sigma PDM.
//
// Three kernels per call, all instance-parallel (no instance ever talks to another):
//   chain_front_kernel  warp = 16 instances x {L, R}: lanes 0-15 carry the left channel, lanes
//                       16-31 the right channel of the same instances; the stereo-linked leveller
//                       and the crossfeed L<->R mix exchange values with __shfl_xor(.., 16).
//                       Writes the two master signals to a frame-major scratch.
//   chain_out_kernel    warp = one output index x 32 instances: matrix mix, 10-band EQ (EqBank,
//                       same register-resident cascade as eq_f32.cu), gain, delay ring in HBM,
//                       peak/clip metering, float -> 24-bit conversion; the sub output leaves
//                       Q28 samples for the modulator.
//   chain_pdm_kernel    one instance per lane: 256x oversampled 2nd-order error-feedback
//                       delta-sigma with the noise-shaped xorshift32 dither, 8 words per frame.
// All per-instance parameters and states are SoA arrays with the instance index innermost, so a
// warp touches consecutive addresses.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "eq_kernels.cuh"
#include "eq_core.cuh"
#include "chain_pdm.cuh"
#include "chain_streams.cuh"

namespace dspi {
namespace {

using namespace core;

constexpr int kOuts = DSPI_CHAIN_OUTPUTS;
constexpr int kRoles = DSPI_CHAIN_EQ_CHANNELS;
constexpr int kMaxDelay = DSPI_CHAIN_MAX_DELAY;
constexpr int kLa = DSPI_LA_SAMPLES;
constexpr int kPkt = DSPI_PACKET_MAX;

enum : uint8_t { F_BYPASS_MASTER = 1, F_LOUD = 2, F_XFEED = 4, F_LEV = 8, F_LOOKAHEAD = 16, F_ANY_DELAY = 32, F_SUB_ON = 64 };
enum : uint8_t { O_ENABLED = 1, O_MUTE = 2, O_PAIR_OFF = 4 };

struct ChainDev {
    uint32_t N, N_pad, nb, max_frames;
    float *coef; uint64_t *modes;                 // packed EQ store, channel = role * N_pad + instance
    float *preamp;                                // [2][N_pad]
    uint8_t *flags;                               // [N_pad] F_*
    float *loud_c; float *loud_st; uint8_t *loud_byp;   // [2 j][6][N_pad], [2 side][2 j][2][N_pad], [N_pad] bit j
    float *xf;                                    // [7][N_pad] lp_a0 lp_b1 lp_L lp_R ap_a ap_L ap_R
    float *lev_c; float *lev_s; uint32_t *lev_idx; float *lev_la;   // [9][N_pad], [5][N_pad], [N_pad], [2][480][N_pad]
    float *o_gl, *o_gr, *o_gain; uint8_t *o_flags; int32_t *o_dly;  // [9][N_pad]
    float *dline; uint32_t *widx_in, *widx_out;   // [9][N_pad][4096], [N_pad]
    int32_t *pdm;                                 // [9][N_pad] err1 err2 x1 x2 y1 y2 err_acc rng fade_in_pos
    uint16_t *peaks; uint16_t *clip;              // [11][N_pad], [N_pad]
    float *master; int32_t *subq;                 // [2][max_frames][N_pad], [max_frames][N_pad]
};

// a*b + c, c - a*b in the flavour's rounding (scalar: negation is free)
template <bool FUSED> __device__ __forceinline__ float fm(float a, float b, float c)
{
    if (FUSED) return __fmaf_rn(a, b, c);
    return __fadd_rn(__fmul_rn(a, b), c);
}
template <bool FUSED> __device__ __forceinline__ float fnm(float a, float b, float c)    // c - a*b
{
    if (FUSED) return __fmaf_rn(-a, b, c);
    return __fadd_rn(c, -__fmul_rn(a, b));
}

__device__ __forceinline__ const float *eq_base(const ChainDev &d, uint32_t role, uint32_t inst)
{
    const uint32_t ch = role * d.N_pad + inst;
    return d.coef + (size_t)(ch >> 5) * kMaxBands * 8 * 32 + (ch & 31);
}

// run one register tile through the bank unless this lane must skip it (state then stays frozen:
// usb_audio.c:721-728 bypass_master_eq, :879-884 muted / disabled outputs)
template <bool FUSED, int NB>
__device__ __forceinline__ void bank_run_masked(EqBank<float, FUSED, NB> &bank, float (&x)[kSub], int nvalid, bool skip)
{
    if (!__any_sync(0xffffffffu, skip)) {
        bank.run(x, nvalid, 0.0f);
        return;
    }
    float keep_x[kSub], keep_s[NB][2];
#pragma unroll
    for (int i = 0; i < kSub; i++) keep_x[i] = x[i];
#pragma unroll
    for (int b = 0; b < NB; b++) { keep_s[b][0] = bank.st[b][0]; keep_s[b][1] = bank.st[b][1]; }
    bank.run(x, nvalid, 0.0f);
    if (skip) {
#pragma unroll
        for (int i = 0; i < kSub; i++) x[i] = keep_x[i];
#pragma unroll
        for (int b = 0; b < NB; b++) { bank.st[b][0] = keep_s[b][0]; bank.st[b][1] = keep_s[b][1]; }
    }
}

// leveller.c:124-139
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
// front: unpack + preamp, loudness, master EQ, leveller, crossfeed, input peaks
// ---------------------------------------------------------------------------------------------
template <bool FUSED, int NB>
__global__ void __launch_bounds__(128, 1)
chain_front_kernel(ChainDev d, const uint8_t *__restrict__ pcm, uint32_t bit_depth, uint32_t p0, uint32_t n_packets, uint32_t fpp, uint32_t F)
{
    extern __shared__ float smem[];                       // [warps][kPkt][32] packet column + the same again for look-ahead reads
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t side = lane >> 4;
    const uint32_t inst = (blockIdx.x * (blockDim.x >> 5) + warp) * 16 + (lane & 15);
    if ((blockIdx.x * (blockDim.x >> 5) + warp) * 16 >= d.N_pad) return;
    const bool live = inst < d.N;
    float *xs = smem + (size_t)warp * kPkt * 32 + lane;   // xs[t * 32]
    float *hs = smem + (size_t)(blockDim.x >> 5) * kPkt * 32 + (size_t)warp * kPkt * 32 + lane;   // held look-ahead samples
    const uint32_t Np = d.N_pad;

    const uint8_t flags = d.flags[inst];
    const bool loud_on = flags & F_LOUD, lev_on = flags & F_LEV, xf_on = flags & F_XFEED;
    const bool skip_master = flags & F_BYPASS_MASTER;
    const bool lookahead = flags & F_LOOKAHEAD;
    const float preamp = d.preamp[side * Np + inst];

    EqBank<float, FUSED, NB> bank;
    float *my_coef = const_cast<float *>(eq_base(d, side, inst));
    {
        const uint64_t *mp[1] = { d.modes + side * Np + inst };
        bank.load(my_coef, mp, d.nb);
    }
    // loudness: 2 general-mix SVF shelves per side (usb_audio.c:689-718)
    float lc[2][6], ls[2][2];
    const uint8_t loud_byp = d.loud_byp[inst];
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int k = 0; k < 6; k++) lc[j][k] = d.loud_c[(j * 6 + k) * Np + inst];
        ls[j][0] = d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst];
        ls[j][1] = d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst];
    }
    // crossfeed (crossfeed.c:132-156): this lane owns its side's lowpass / all-pass state
    const float xf_a0 = d.xf[0 * Np + inst], xf_b1 = d.xf[1 * Np + inst], xf_ap = d.xf[4 * Np + inst];
    float xf_lp = d.xf[(2 + side) * Np + inst], xf_as = d.xf[(5 + side) * Np + inst];
    // leveller (leveller.c:148-262)
    float lvc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) lvc[k] = d.lev_c[k * Np + inst];
    float env = d.lev_s[side * Np + inst];
    float smooth_db = d.lev_s[2 * Np + inst], gain_lin = d.lev_s[3 * Np + inst], gain_prev = d.lev_s[4 * Np + inst];
    uint32_t la_idx = d.lev_idx[inst];
    float *la_buf = d.lev_la + (size_t)side * kLa * Np + inst;

    const uint32_t bpf = bit_depth == 24 ? 6u : 4u;
    const uint8_t *my_pcm = pcm + ((size_t)inst * F) * bpf + side * (bpf / 2);
    float gain_in;
    if (bit_depth == 24) gain_in = __fmul_rn(1.0f / 8388608.0f, preamp);     // usb_audio.c:601-603
    else gain_in = __fmul_rn(1.0f / 32768.0f, preamp);                       // :680-681

    float peak_in = 0.0f;
    uint16_t clip = 0;
    const uint32_t f_end = (p0 + n_packets) * fpp;
    for (uint32_t p = p0; p < p0 + n_packets; p++) {
        const uint32_t f0 = p * fpp;
        // The leveller's look-ahead ring is read one slot per sample, each read just before that slot is
        // overwritten (leveller.c:231-237), and a packet (<= 192 frames) never laps the 480-slot ring:
        // all of this packet's reads can be issued now, as asynchronous copies into shared memory that
        // complete behind passes 1-2, instead of one exposed HBM round trip per sample.
        if (lev_on && lookahead) {
            uint32_t idx = la_idx;
            for (uint32_t i = 0; i < fpp; i++) {
                cp_async_4(hs + i * 32, la_buf + (size_t)idx * Np);
                if (++idx >= (uint32_t)kLa) idx = 0;
            }
        }
        cp_async_commit();
        // ---- PASS 1 + loudness + PASS 2 (master EQ), register tiles of 8 ----
        for (uint32_t t0 = 0; t0 < fpp; t0 += kSub) {
            const int nvalid = min((int)kSub, (int)(fpp - t0));
            float x[kSub];
            if (live) {                                    // next tile's PCM bytes (a private 6 B/frame stream per lane) towards L1
                const uint32_t fn = f0 + t0 + kSub;
                if (fn < f_end) {
                    prefetch_l1(my_pcm + (size_t)fn * bpf);
                    prefetch_l1(my_pcm + (size_t)(min(fn + (uint32_t)kSub, f_end) - 1) * bpf);
                }
            }
#pragma unroll
            for (int i = 0; i < kSub; i++) {
                int32_t s = 0;
                if (live && i < nvalid) {
                    const uint8_t *q = my_pcm + (size_t)(f0 + t0 + i) * bpf;
                    if (bit_depth == 24) s = ((int32_t)((uint32_t)q[2] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[0] << 8)) >> 8;
                    else s = (int16_t)((uint16_t)q[0] | (uint16_t)q[1] << 8);
                }
                x[i] = __fmul_rn((float)s, gain_in);                         // :645-648 / :683-684
            }
            if (loud_on) {
#pragma unroll
                for (int i = 0; i < kSub; i++) {
                    if (i < nvalid) {
                        float v = x[i];
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            if ((loud_byp >> j) & 1) continue;
                            const float v3 = __fadd_rn(v, -ls[j][1]);
                            const float pp = __fmul_rn(lc[j][1], v3);
                            float t, v1, v2;
                            if (FUSED) {
                                t = __fmaf_rn(lc[j][1], ls[j][0], ls[j][1]);
                                v1 = __fmaf_rn(lc[j][0], ls[j][0], pp);
                                v2 = __fmaf_rn(lc[j][2], v3, t);
                            } else {
                                t = __fadd_rn(ls[j][1], __fmul_rn(lc[j][1], ls[j][0]));
                                v1 = __fadd_rn(__fmul_rn(lc[j][0], ls[j][0]), pp);
                                v2 = __fadd_rn(t, __fmul_rn(lc[j][2], v3));
                            }
                            ls[j][0] = __fmaf_rn(2.0f, v1, -ls[j][0]);
                            ls[j][1] = __fmaf_rn(2.0f, v2, -ls[j][1]);
                            v = fm<FUSED>(lc[j][5], v2, fm<FUSED>(lc[j][3], v, __fmul_rn(lc[j][4], v1)));   // :702
                        }
                        x[i] = v;
                    }
                }
            }
            bank_run_masked<FUSED, NB>(bank, x, nvalid, skip_master);
#pragma unroll
            for (int i = 0; i < kSub; i++)
                if (i < nvalid) xs[(t0 + i) * 32] = x[i];
        }
        __syncwarp();

        // ---- PASS 2.5: leveller ----
        if (__any_sync(0xffffffffu, lev_on)) {
            const float a_rms = lvc[0], one_minus = __fadd_rn(1.0f, -a_rms);
            float e = env;
            for (uint32_t i = 0; i < fpp; i++) {                             // leveller.c:161-166
                const float s = xs[i * 32];
                e = fm<FUSED>(a_rms, e, __fmul_rn(one_minus, __fmul_rn(s, s)));
            }
            if (e < 1e-30f) e = 0.0f;                                        // :169-170
            const float e_other = __shfl_xor_sync(0xffffffffu, e, 16);
            const float env_l = side ? e_other : e, env_r = side ? e : e_other;
            const float rms_sq = (env_l > env_r) ? env_l : env_r;            // :177
            // per-block libm: evaluated in double and rounded once (DESIGN.md "libm policy")
            const float rms_db = __fmul_rn(10.0f, (float)log10((double)__fadd_rn(rms_sq, 1e-30f)));
            float gc_db;
            if (rms_db < lvc[7]) gc_db = 0.0f;
            else {
                gc_db = gain_computer(rms_db, lvc[3], lvc[4], lvc[5]);
                gc_db = __fadd_rn(gc_db, lvc[6]);
                if (gc_db > lvc[8]) gc_db = lvc[8];
            }
            const float alpha_s = (gc_db < smooth_db) ? lvc[1] : lvc[2];     // :198
            const float alpha = (float)pow((double)alpha_s, (double)(float)fpp);
            const float new_smooth = fm<FUSED>(alpha, smooth_db, __fmul_rn(__fadd_rn(1.0f, -alpha), gc_db));
            const float new_gain = (float)pow(10.0, (double)__fdiv_rn(new_smooth, 20.0f));
            // every lane walks the same shuffles; only instances with the leveller on commit results
            const float prev_for_ramp = gain_lin;
            float gain, gain_step;
            if (fpp == 1) { gain = new_gain; gain_step = 0.0f; }
            else { gain_step = __fdiv_rn(__fadd_rn(new_gain, -prev_for_ramp), (float)(fpp - 1)); gain = prev_for_ramp; }
            cp_async_wait_all();                                             // this lane's look-ahead reads have landed
            for (uint32_t i = 0; i < fpp; i++) {                             // :228-259
                float o = xs[i * 32];
                if (lev_on && lookahead) {
                    const float held = hs[i * 32];
                    la_buf[(size_t)la_idx * Np] = o;
                    o = held;
                    la_idx++;
                    if (la_idx >= (uint32_t)kLa) la_idx = 0;
                }
                const float ao = fabsf(o);
                const float ao_other = __shfl_xor_sync(0xffffffffu, ao, 16);
                const float al = side ? ao_other : ao, ar = side ? ao : ao_other;
                float peak = al;
                if (ar > peak) peak = ar;
                float g = gain;
                if (peak > 0.0f && g > 1.0f) {
                    const float max_g = __fdiv_rn(0.70795f, peak);
                    if (max_g < g) g = (max_g > 1.0f) ? max_g : 1.0f;
                }
                if (lev_on) xs[i * 32] = __fmul_rn(o, g);
                gain = __fadd_rn(gain, gain_step);
            }
            if (lev_on) {
                env = e;
                smooth_db = new_smooth;
                gain_prev = gain_lin;
                gain_lin = new_gain;
            }
        }

        // ---- PASS 3: input peaks, then crossfeed (usb_audio.c:741-749) ----
        float pk = 0.0f;
        float *mout = d.master + ((size_t)side * d.max_frames + f0) * Np + inst;
        for (uint32_t i = 0; i < fpp; i++) {
            float v = xs[i * 32];
            const float a = fabsf(v);
            if (a > pk) pk = a;
            float lp = 0.0f, ap = 0.0f;
            if (xf_on) {
                lp = fm<FUSED>(xf_a0, v, __fmul_rn(xf_b1, xf_lp));           // crossfeed.c:137-138
                xf_lp = lp;
                ap = fm<FUSED>(xf_ap, lp, xf_as);                            // :146 / :148
                xf_as = fnm<FUSED>(xf_ap, ap, lp);                           // :147 / :149
            }
            const float ap_other = __shfl_xor_sync(0xffffffffu, ap, 16);
            if (xf_on) v = __fadd_rn(__fadd_rn(v, -lp), ap_other);           // :154-155
            mout[(size_t)i * Np] = v;
        }
        peak_in = pk;                                                        // peaks describe the last packet
        if (pk > 1.001f) clip |= (uint16_t)(1u << side);                     // config.h:53
        __syncwarp();
    }

    // ---- state back ----
    bank.store(my_coef);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst] = ls[j][0];
        d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst] = ls[j][1];
    }
    d.xf[(2 + side) * Np + inst] = xf_lp;
    d.xf[(5 + side) * Np + inst] = xf_as;
    d.lev_s[side * Np + inst] = env;
    if (side == 0) {
        d.lev_s[2 * Np + inst] = smooth_db;
        d.lev_s[3 * Np + inst] = gain_lin;
        d.lev_s[4 * Np + inst] = gain_prev;
        d.lev_idx[inst] = la_idx;
    }
    d.peaks[side * Np + inst] = (uint16_t)__fmul_rn(fminf(1.0f, peak_in), 32767.0f);    // usb_audio.c:963-964
    const uint16_t clip_other = (uint16_t)__shfl_xor_sync(0xffffffffu, (uint32_t)clip, 16);
    if (side == 0) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), (unsigned int)(clip | clip_other) << (16 * (inst & 1)));
}

// ---------------------------------------------------------------------------------------------
// outputs: matrix, per-output EQ, gain, delay, peaks, 24-bit conversion / Q28 for the modulator
// ---------------------------------------------------------------------------------------------
template <bool FUSED, int NB>
__global__ void __launch_bounds__(128, 3)
chain_out_kernel(ChainDev d, uint32_t p0, uint32_t n_packets, uint32_t fpp, uint32_t F, int32_t *__restrict__ spdif_out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + warp;
    const uint32_t groups = d.N_pad / 32;
    if (wid >= groups * kOuts) return;
    const uint32_t o = wid / groups;                         // output index: warp-uniform
    const uint32_t inst = (wid % groups) * 32 + lane;
    const bool live = inst < d.N;
    const uint32_t Np = d.N_pad;

    const uint8_t of = d.o_flags[o * Np + inst];
    const bool enabled = of & O_ENABLED, mute = of & O_MUTE, pair_off = of & O_PAIR_OFF;
    const float gl = d.o_gl[o * Np + inst], gr = d.o_gr[o * Np + inst], gain = d.o_gain[o * Np + inst];
    const int32_t dly = d.o_dly[o * Np + inst];
    const bool delay_on = (d.flags[inst] & F_ANY_DELAY) && dly > 0;          // usb_audio.c:898-901
    const bool any_delay = d.flags[inst] & F_ANY_DELAY;
    const bool ring_early = delay_on && dly >= kSub && dly < kMaxDelay;
    const uint32_t f_end = (p0 + n_packets) * fpp;
    uint32_t widx = d.widx_in[inst];
    float *ring = d.dline + ((size_t)o * Np + inst) * kMaxDelay;     // one contiguous ring per (output, instance)

    EqBank<float, FUSED, NB> bank;
    float *my_coef = const_cast<float *>(eq_base(d, 2 + o, inst));
    {
        const uint64_t *mp[1] = { d.modes + (2 + o) * Np + inst };
        bank.load(my_coef, mp, d.nb);
    }
    const bool skip_eq = !enabled || mute;                                   // :878-884
    const int mixcase = !enabled ? 0 : (gl != 0.0f && gr != 0.0f) ? 3 : (gl != 0.0f) ? 1 : (gr != 0.0f) ? 2 : 0;   // :767-778

    // The master L/R rows a tile needs are staged through shared memory by asynchronous copies issued two
    // tiles ahead (each lane fetches its own instance's samples, so no warp synchronisation is needed):
    // the load latency (L2 / HBM) overlaps the EQ arithmetic of the tiles in between.
    constexpr int kLrStages = 3;
    __shared__ float lr_stage[4][kLrStages][2][kSub][32];
    float *lr = &lr_stage[warp][0][0][0][lane];
    const uint32_t tpp = (fpp + kSub - 1) / kSub, n_tiles = n_packets * tpp;
    auto issue_lr = [&](uint32_t n) {
        if (n < n_tiles) {
            const uint32_t pn = p0 + n / tpp, tt = (n % tpp) * kSub;
            float *dst = lr + (n % kLrStages) * (2 * kSub * 32);
            const float *src = d.master + ((size_t)pn * fpp + tt) * Np + inst;
#pragma unroll
            for (int i = 0; i < kSub; i++) {
                if (tt + i < fpp) {
                    if (mixcase & 1) cp_async_4(dst + i * 32, src + (size_t)i * Np);
                    if (mixcase & 2) cp_async_4(dst + (kSub + i) * 32, src + ((size_t)d.max_frames + i) * Np);
                }
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int n = 0; n < kLrStages - 1; n++) issue_lr(n);
    uint32_t tile_n = 0;

    float peak_last = 0.0f;
    uint16_t clip = 0;
    const bool is_sub = o == kOuts - 1;
    int32_t *my_spdif = nullptr;
    if (!is_sub && spdif_out && live) my_spdif = spdif_out + (((size_t)inst * 4 + (o >> 1)) * F) * 2 + (o & 1);

    for (uint32_t p = p0; p < p0 + n_packets; p++) {
        const uint32_t f0 = p * fpp;
        float pk = 0.0f;
        uint32_t w = widx;
        for (uint32_t t0 = 0; t0 < fpp; t0 += kSub) {
            const int nvalid = min((int)kSub, (int)(fpp - t0));
            float x[kSub];
            issue_lr(tile_n + kLrStages - 1);                                // master L/R of the tile two ahead
            cp_async_wait<kLrStages - 1>();                                  // ... and this tile's have landed
            const float *cur = lr + (tile_n % kLrStages) * (2 * kSub * 32);
            tile_n++;
            // Delay ring (:902-909 is write-then-read per sample).  For kSub <= dly < MAX the slots this
            // tile reads were written by earlier tiles (and the slots it writes are not read in it), so
            // the reads are issued NOW and their HBM latency hides behind the matrix + EQ arithmetic.
            float rd[kSub];
            if (ring_early) {
#pragma unroll
                for (int i = 0; i < kSub; i++)
                    if (i < nvalid) rd[i] = ring[(w + i - (uint32_t)dly) & (kMaxDelay - 1)];
            }
#pragma unroll
            for (int i = 0; i < kSub; i++) {
                float l = 0.0f, r = 0.0f;
                if (i < nvalid) {
                    if (mixcase & 1) l = cur[i * 32];
                    if (mixcase & 2) r = cur[(kSub + i) * 32];
                }
                float v;
                if (mixcase == 3) v = fm<FUSED>(l, gl, __fmul_rn(r, gr));    // :769
                else if (mixcase == 1) v = __fmul_rn(l, gl);
                else if (mixcase == 2) v = __fmul_rn(r, gr);
                else v = 0.0f;
                x[i] = v;
            }
            bank_run_masked<FUSED, NB>(bank, x, nvalid, skip_eq);
#pragma unroll
            for (int i = 0; i < kSub; i++) {
                if (enabled) {                                               // :885-894
                    if (gain == 0.0f) x[i] = 0.0f;
                    else if (gain != 1.0f) x[i] = __fmul_rn(x[i], gain);
                }
            }
            if (ring_early) {
#pragma unroll
                for (int i = 0; i < kSub; i++)
                    if (i < nvalid) { ring[(w + i) & (kMaxDelay - 1)] = x[i]; x[i] = rd[i]; }
            } else if (delay_on) {                                           // dly < kSub, or dly == MAX (aliases to 0, SURVEY a-10)
                for (int i = 0; i < nvalid; i++) {
                    ring[(w + i) & (kMaxDelay - 1)] = x[i];
                    x[i] = ring[(w + i - (uint32_t)dly) & (kMaxDelay - 1)];
                }
            }
            w = (w + nvalid) & (kMaxDelay - 1);
#pragma unroll
            for (int i = 0; i < kSub; i++) {
                if (i >= nvalid) break;
                const float v = x[i];
                const float a = fabsf(v);
                if (a > pk) pk = a;
                if (is_sub) {
                    if (enabled) d.subq[(size_t)(f0 + t0 + i) * Np + inst] = __float2int_rz(__fmul_rn(v, 268435456.0f));   // :953 (saturating)
                } else if (my_spdif) {
                    int32_t word = 0;
                    if (!pair_off) {
                        const float c = fmaxf(-1.0f, fminf(1.0f, v));       // :936-939
                        word = __float2int_rz(__fmul_rn(c, 8388607.0f));
                    }
                    my_spdif[(size_t)(f0 + t0 + i) * 2] = word;
                }
            }
        }
        if (any_delay) widx = (widx + fpp) & (kMaxDelay - 1);               // :911
        peak_last = pk;
        if (pk > 1.001f && (!is_sub || enabled)) clip |= 1;
    }
    bank.store(my_coef);
    uint16_t pq = (uint16_t)__fmul_rn(fminf(1.0f, peak_last), 32767.0f);    // :921 / :950
    if (is_sub && !enabled) pq = 0;                                          // :957
    d.peaks[(2 + o) * Np + inst] = pq;
    if (clip) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), (1u << (2 + o)) << (16 * (inst & 1)));
    if (o == 0) d.widx_out[inst] = widx;
}

// ---------------------------------------------------------------------------------------------
// delta-sigma PDM (chain_pdm.cuh): one instance per lane
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
chain_pdm_kernel(ChainDev d, uint32_t f_begin, uint32_t f_end, uint32_t F, uint32_t *__restrict__ pdm_out)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= d.N) return;
    if (!(d.flags[inst] & F_SUB_ON)) return;                                 // usb_audio.c:944
    pdm_modulate_frames(d.pdm, d.subq, d.N_pad, inst, f_begin, f_end, F, pdm_out);
}

// filters[][] of n instances (instance-major AoS) -> packed store with channel = role * N_pad + inst
__global__ void chain_pack_kernel(const dspi_biquad_f32 *__restrict__ aos, uint32_t inst0, uint32_t n, ChainDev d)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kRoles) return;
    const uint32_t inst = inst0 + i / kRoles, role = i % kRoles;
    const uint32_t ch = role * d.N_pad + inst, g = ch >> 5, lane = ch & 31;
    uint64_t mw = 0;
    for (int b = 0; b < kMaxBands; b++) {
        const dspi_biquad_f32 &q = aos[((size_t)inst * kRoles + role) * kMaxBands + b];
        float v[8];
        uint32_t mode;
        if (q.bypass) mode = kModeBypass;
        else if (!q.use_svf) mode = kModeTdf2;
        else mode = q.svf_type == DSPI_FILTER_LOWPASS ? kModeSvfLP : q.svf_type == DSPI_FILTER_HIGHPASS ? kModeSvfHP
                  : q.svf_type == DSPI_FILTER_PEAKING ? kModeSvfPK : kModeSvfSH;
        if (q.use_svf && !q.bypass) {
            v[0] = q.sva1; v[1] = q.sva2; v[2] = q.sva3; v[3] = q.svm0; v[4] = q.svm1; v[5] = q.svm2; v[6] = q.svic1eq; v[7] = q.svic2eq;
        } else {
            v[0] = q.b0; v[1] = q.b1; v[2] = q.b2; v[3] = -q.a1; v[4] = -q.a2; v[5] = 0.0f; v[6] = q.s1; v[7] = q.s2;
        }
        mw |= (uint64_t)mode << (4 * b);
        for (int k = 0; k < 8; k++) d.coef[(((size_t)g * kMaxBands + b) * 8 + k) * 32 + lane] = v[k];
    }
    d.modes[ch] = mw;
}

__global__ void chain_unpack_kernel(dspi_biquad_f32 *__restrict__ aos, uint32_t inst0, uint32_t n, ChainDev d)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kRoles) return;
    const uint32_t inst = inst0 + i / kRoles, role = i % kRoles;
    const uint32_t ch = role * d.N_pad + inst, g = ch >> 5, lane = ch & 31;
    for (int b = 0; b < kMaxBands; b++) {
        dspi_biquad_f32 &q = aos[((size_t)inst * kRoles + role) * kMaxBands + b];
        if (q.bypass) continue;
        const float s0 = d.coef[(((size_t)g * kMaxBands + b) * 8 + 6) * 32 + lane];
        const float s1 = d.coef[(((size_t)g * kMaxBands + b) * 8 + 7) * 32 + lane];
        if (q.use_svf) { q.svic1eq = s0; q.svic2eq = s1; }
        else { q.s1 = s0; q.s2 = s1; }
    }
}

__global__ void chain_status_kernel(ChainDev d, dspi_status *__restrict__ out)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= d.N) return;
    dspi_status s;
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

}  // namespace
}  // namespace dspi

using dspi::ChainDev;
using dspi::fail;

#define CU_OK(expr)                                                                                         \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) return fail(DSPI_ECUDA, "%s -> %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

struct dspi_chain {
    dspi_chain_desc desc;
    ChainDev d;
    cudaStream_t stream;                  // the engine stream callers see; stages run on st.* between ev_begin and ev_done
    dspi::ChainStreams st;
    dspi_biquad_f32 *d_aos;          // [N_pad][11][12] instance-major mirror of filters[][]
    std::vector<void *> allocs;
    uint64_t launches;
    void *d_pcm; size_t pcm_bytes;   // host-path staging
    int32_t *d_spdif; size_t spdif_bytes;
    uint32_t *d_pdmout; size_t pdmout_bytes;
    dspi_status *d_status;
};

namespace {

template <typename T>
cudaError_t dev_alloc(dspi_chain *c, T **p, size_t count, bool zero = true)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return e;
    c->allocs.push_back(q);
    *p = (T *)q;
    return zero ? cudaMemsetAsync(q, 0, count * sizeof(T), c->stream) : cudaSuccess;
}

// state that leveller_reset_state() / the PDM restart path define as non-zero
cudaError_t init_states(dspi_chain *c)
{
    const uint32_t Np = c->d.N_pad;
    std::vector<float> one(Np, 1.0f);
    std::vector<int32_t> seed(Np, 123456789);                               // pdm_generator.c:62
    cudaError_t e;
    if ((e = cudaMemsetAsync(c->d.lev_s, 0, (size_t)5 * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(c->d.lev_s + 3 * Np, one.data(), Np * 4, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;   // gain_linear = 1
    if ((e = cudaMemcpyAsync(c->d.lev_s + 4 * Np, one.data(), Np * 4, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;   // gain_prev_linear = 1
    if ((e = cudaMemsetAsync(c->d.lev_idx, 0, (size_t)Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.lev_la, 0, (size_t)2 * dspi::kLa * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.loud_st, 0, (size_t)8 * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.dline, 0, (size_t)dspi::kOuts * dspi::kMaxDelay * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.widx_in, 0, (size_t)Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.widx_out, 0, (size_t)Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.pdm, 0, (size_t)9 * Np * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(c->d.pdm + 7 * Np, seed.data(), Np * 4, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.peaks, 0, (size_t)dspi::kRoles * Np * 2, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(c->d.clip, 0, (size_t)Np * 2, c->stream)) != cudaSuccess) return e;
    return cudaStreamSynchronize(c->stream);
}

template <bool FUSED>
int launch_chain(dspi_chain *c, const void *d_pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t fpp, int32_t *d_spdif, uint32_t *d_pdm,
                 dspi_status *d_status)
{
    const uint32_t F = n_packets * fpp;
    auto front = dspi::chain_front_kernel<FUSED, 10>;
    const size_t smem = (size_t)2 * 4 * dspi::kPkt * 32 * 4;      // packet column + look-ahead column per warp
    static bool configured = false;
    if (!configured) { CU_OK(cudaFuncSetAttribute(front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; }
    // Stage pipeline over packet slices on three streams (chain_streams.cuh).
    dspi::ChainStreams &st = c->st;
    const uint32_t n_slices = n_packets < (uint32_t)dspi::ChainStreams::kMaxSlices ? n_packets : (uint32_t)dspi::ChainStreams::kMaxSlices;
    CU_OK(cudaEventRecord(st.ev_begin, c->stream));
    CU_OK(cudaStreamWaitEvent(st.s_front, st.ev_begin, 0));
    for (uint32_t sl = 0; sl < n_slices; sl++) {
        const uint32_t p0 = (uint32_t)((uint64_t)n_packets * sl / n_slices), p1 = (uint32_t)((uint64_t)n_packets * (sl + 1) / n_slices);
        const ChainDev d = c->d;
        const uint32_t fwarps = d.N_pad / 16, owarps = d.N_pad / 32 * dspi::kOuts;
        front<<<(fwarps + 3) / 4, 128, smem, st.s_front>>>(d, (const uint8_t *)d_pcm, bit_depth, p0, p1 - p0, fpp, F);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(st.ev_front[sl], st.s_front));
        CU_OK(cudaStreamWaitEvent(st.s_out, st.ev_front[sl], 0));
        dspi::chain_out_kernel<FUSED, 10><<<(owarps + 3) / 4, 128, 0, st.s_out>>>(d, p0, p1 - p0, fpp, F, d_spdif);
        CU_OK(cudaGetLastError());
        std::swap(c->d.widx_in, c->d.widx_out);
        CU_OK(cudaEventRecord(st.ev_out[sl], st.s_out));
        CU_OK(cudaStreamWaitEvent(st.s_pdm, st.ev_out[sl], 0));
        dspi::chain_pdm_kernel<<<(d.N + 63) / 64, 64, 0, st.s_pdm>>>(d, p0 * fpp, p1 * fpp, F, d_pdm);
        CU_OK(cudaGetLastError());
        c->launches += 3;
    }
    // the last modulator launch is ordered after every other stage launch of this call
    CU_OK(cudaEventRecord(st.ev_done, st.s_pdm));
    CU_OK(cudaStreamWaitEvent(c->stream, st.ev_done, 0));         // later work on the engine stream sees all outputs
    if (d_status) {
        dspi::chain_status_kernel<<<(c->d.N + 127) / 128, 128, 0, c->stream>>>(c->d, d_status);
        CU_OK(cudaGetLastError());
        c->launches++;
    }
    return DSPI_OK;
}

}  // namespace

extern "C" {

int32_t dspi_delay_samples(float delay_ms, float sample_rate, int is_last)
{
    if (is_last) delay_ms += (float)128 / sample_rate * 1000.0f;            // SUB_ALIGN_SAMPLES, config.h:93-95
    const float x = delay_ms * sample_rate / 1000.0f;
    int32_t s = (x != x) ? 0 : (x >= 2147483648.0f ? INT32_MAX : (x <= -2147483648.0f ? INT32_MIN : (int32_t)x));
    if (s > DSPI_CHAIN_MAX_DELAY) s = DSPI_CHAIN_MAX_DELAY;
    if (s < 0) s = 0;
    return s;
}

int dspi_chain_destroy(dspi_chain *c)
{
    if (!c) return DSPI_OK;
    cudaSetDevice(c->desc.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->st.destroy();
    for (void *p : c->allocs) cudaFree(p);
    if (c->d_pcm) cudaFree(c->d_pcm);
    if (c->d_spdif) cudaFree(c->d_spdif);
    if (c->d_pdmout) cudaFree(c->d_pdmout);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    cudaGetLastError();
    return DSPI_OK;
}

int dspi_chain_create(dspi_chain **out, const dspi_chain_desc *desc)
{
    if (!out || !desc) return fail(DSPI_EINVAL, "null argument");
    *out = nullptr;
    if (desc->arith != DSPI_ARITH_F32_FUSED && desc->arith != DSPI_ARITH_F32_STRICT) return fail(DSPI_EINVAL, "chain engines are float (arith 0 or 1)");
    if (desc->n_instances == 0 || desc->max_frames == 0) return fail(DSPI_EINVAL, "n_instances and max_frames must be > 0");
    if (desc->n_bands != 10) return fail(DSPI_EINVAL, "chain engines run channel_band_counts = 10");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(DSPI_ENODEV, "no CUDA device (there is no CPU fallback)"); }
    if (desc->device < 0 || desc->device >= ndev) return fail(DSPI_ENODEV, "device %d out of range", desc->device);
    int major = 0;
    CU_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, desc->device));
    if (major != 10) return fail(DSPI_ENODEV, "device %d is not sm_100", desc->device);
    CU_OK(cudaSetDevice(desc->device));
    dspi_chain *c = new (std::nothrow) dspi_chain();
    if (!c) return fail(DSPI_ENOMEM, "host allocation failed");
    c->stream = nullptr;
    c->st = dspi::ChainStreams();
    c->d_aos = nullptr; c->launches = 0; c->d_pcm = nullptr; c->pcm_bytes = 0; c->d_spdif = nullptr; c->spdif_bytes = 0;
    c->d_pdmout = nullptr; c->pdmout_bytes = 0; c->d_status = nullptr;
    c->desc = *desc;
    ChainDev &d = c->d;
    memset(&d, 0, sizeof(d));
    d.N = desc->n_instances;
    d.N_pad = (d.N + 31) / 32 * 32;
    d.nb = desc->n_bands;
    d.max_frames = desc->max_frames;
    const size_t Np = d.N_pad;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = c->st.create();
#define TRY(x) if (e == cudaSuccess) e = (x)
    TRY(dev_alloc(c, &c->d_aos, Np * dspi::kRoles * DSPI_MAX_BANDS));
    TRY(dev_alloc(c, &d.coef, Np * dspi::kRoles * DSPI_MAX_BANDS * 8));
    TRY(dev_alloc(c, &d.modes, Np * dspi::kRoles));
    TRY(dev_alloc(c, &d.preamp, 2 * Np));
    TRY(dev_alloc(c, &d.flags, Np));
    TRY(dev_alloc(c, &d.loud_c, 12 * Np));
    TRY(dev_alloc(c, &d.loud_st, 8 * Np));
    TRY(dev_alloc(c, &d.loud_byp, Np));
    TRY(dev_alloc(c, &d.xf, 7 * Np));
    TRY(dev_alloc(c, &d.lev_c, 9 * Np));
    TRY(dev_alloc(c, &d.lev_s, 5 * Np));
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
    TRY(dev_alloc(c, &d.master, (size_t)2 * d.max_frames * Np, false));
    TRY(dev_alloc(c, &d.subq, (size_t)d.max_frames * Np, false));
    TRY(dev_alloc(c, &c->d_status, Np));
    TRY(init_states(c));
#undef TRY
    if (e != cudaSuccess) {
        fail(e == cudaErrorMemoryAllocation ? DSPI_ENOMEM : DSPI_ECUDA, "chain setup: %s", cudaGetErrorString(e));
        dspi_chain_destroy(c);
        return e == cudaErrorMemoryAllocation ? DSPI_ENOMEM : DSPI_ECUDA;
    }
    *out = c;
    return DSPI_OK;
}

int dspi_chain_reset_state(dspi_chain *c)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(init_states(c));
    return DSPI_OK;
}

int dspi_chain_set_params(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_chain_params_f32 *params)
{
    if (!c || !params) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const ChainDev &d = c->d;
    const size_t Np = d.N_pad;
    std::vector<float> preamp(2 * n), loud_c(12 * n), xf(7 * n), lev_c(9 * n), gl(9 * n), gr(9 * n), gain(9 * n);
    std::vector<uint8_t> flags(n), loud_byp(n), oflags(9 * n);
    std::vector<int32_t> dly(9 * n);
    for (uint32_t i = 0; i < n; i++) {
        const dspi_chain_params_f32 &p = params[i];
        // usb_audio.c:569-571
        float vol_mul = p.host_mute ? 0.0f : (float)p.host_vol_mul * (1.0f / 32768.0f);
        vol_mul *= p.preset_mute_gain;
        const float vol_mul_master = vol_mul * p.master_volume_linear;
        preamp[0 * n + i] = p.preamp_linear[0];
        preamp[1 * n + i] = p.preamp_linear[1];
        bool any_delay = false;
        for (int o = 0; o < dspi::kOuts; o++) {
            const dspi_output_channel &oc = p.matrix.outputs[o];
            const dspi_matrix_crosspoint &xl = p.matrix.crosspoints[0][o], &xr = p.matrix.crosspoints[1][o];
            float a = 0.0f, b = 0.0f;                                        // :760-764
            if (xl.enabled) a = xl.phase_invert ? -xl.gain_linear : xl.gain_linear;
            if (xr.enabled) b = xr.phase_invert ? -xr.gain_linear : xr.gain_linear;
            gl[o * n + i] = a;
            gr[o * n + i] = b;
            gain[o * n + i] = oc.mute ? 0.0f : oc.gain_linear * vol_mul_master;    // :886-887
            uint8_t f = (oc.enabled ? dspi::O_ENABLED : 0) | (oc.mute ? dspi::O_MUTE : 0);
            if (o < dspi::kOuts - 1) {
                const int partner = o ^ 1;
                if (!oc.enabled && !p.matrix.outputs[partner].enabled) f |= dspi::O_PAIR_OFF;    // :930-933
            }
            oflags[o * n + i] = f;
            int32_t ds = oc.delay_samples;
            if (ds > DSPI_CHAIN_MAX_DELAY) ds = DSPI_CHAIN_MAX_DELAY;
            if (ds < 0) ds = 0;
            dly[o * n + i] = ds;
            if (ds > 0) any_delay = true;                                    // dsp_pipeline.c:237
        }
        flags[i] = (p.bypass_master_eq ? dspi::F_BYPASS_MASTER : 0) | (p.loudness_enabled ? dspi::F_LOUD : 0) |
                   (p.crossfeed_enabled ? dspi::F_XFEED : 0) | (p.leveller_enabled ? dspi::F_LEV : 0) |
                   (p.leveller_lookahead ? dspi::F_LOOKAHEAD : 0) | (any_delay ? dspi::F_ANY_DELAY : 0) |
                   (p.matrix.outputs[dspi::kOuts - 1].enabled ? dspi::F_SUB_ON : 0);
        loud_byp[i] = (p.loudness[0].bypass ? 1 : 0) | (p.loudness[1].bypass ? 2 : 0);
        for (int j = 0; j < 2; j++) {
            const float v[6] = { p.loudness[j].sva1, p.loudness[j].sva2, p.loudness[j].sva3, p.loudness[j].svm0, p.loudness[j].svm1, p.loudness[j].svm2 };
            for (int k = 0; k < 6; k++) loud_c[(j * 6 + k) * n + i] = v[k];
        }
        const float xv[7] = { p.crossfeed.lp_a0, p.crossfeed.lp_b1, p.crossfeed.lp_state_L, p.crossfeed.lp_state_R,
                              p.crossfeed.ap_a, p.crossfeed.ap_state_L, p.crossfeed.ap_state_R };
        for (int k = 0; k < 7; k++) xf[k * n + i] = xv[k];
        const float *lv = &p.leveller.alpha_rms;
        for (int k = 0; k < 9; k++) lev_c[k * n + i] = lv[k];
    }
    auto put = [&](void *dst_base, const void *src, int rows, size_t elem) -> cudaError_t {
        return cudaMemcpy2DAsync((char *)dst_base + (size_t)inst0 * elem, Np * elem, src, (size_t)n * elem, (size_t)n * elem, rows,
                                 cudaMemcpyHostToDevice, c->stream);
    };
    CU_OK(put(d.preamp, preamp.data(), 2, 4));
    CU_OK(put(d.flags, flags.data(), 1, 1));
    CU_OK(put(d.loud_c, loud_c.data(), 12, 4));
    CU_OK(put(d.loud_byp, loud_byp.data(), 1, 1));
    CU_OK(put(d.xf, xf.data(), 7, 4));
    CU_OK(put(d.lev_c, lev_c.data(), 9, 4));
    CU_OK(put(d.o_gl, gl.data(), 9, 4));
    CU_OK(put(d.o_gr, gr.data(), 9, 4));
    CU_OK(put(d.o_gain, gain.data(), 9, 4));
    CU_OK(put(d.o_flags, oflags.data(), 9, 1));
    CU_OK(put(d.o_dly, dly.data(), 9, 4));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chain_upload_biquads(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_biquad_f32 *biquads)
{
    if (!c || !biquads) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t row = (size_t)dspi::kRoles * DSPI_MAX_BANDS;
    CU_OK(cudaMemcpyAsync(c->d_aos + inst0 * row, biquads, n * row * sizeof(dspi_biquad_f32), cudaMemcpyHostToDevice, c->stream));
    dspi::chain_pack_kernel<<<(n * dspi::kRoles + 127) / 128, 128, 0, c->stream>>>(c->d_aos, inst0, n, c->d);
    CU_OK(cudaGetLastError());
    c->launches++;
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chain_download_biquads(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_biquad_f32 *biquads)
{
    if (!c || !biquads) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    dspi::chain_unpack_kernel<<<(n * dspi::kRoles + 127) / 128, 128, 0, c->stream>>>(c->d_aos, inst0, n, c->d);
    CU_OK(cudaGetLastError());
    c->launches++;
    const size_t row = (size_t)dspi::kRoles * DSPI_MAX_BANDS;
    CU_OK(cudaMemcpyAsync(biquads, c->d_aos + inst0 * row, n * row * sizeof(dspi_biquad_f32), cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

static int check_process(dspi_chain *c, const void *pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t fpp)
{
    if (!c || !pcm) return fail(DSPI_EINVAL, "null argument");
    if (bit_depth != 16 && bit_depth != 24) return fail(DSPI_EINVAL, "bit_depth must be 16 or 24");
    if (fpp == 0 || fpp > DSPI_PACKET_MAX) return fail(DSPI_EINVAL, "frames_per_packet must be 1..%d", DSPI_PACKET_MAX);
    if (n_packets == 0) return fail(DSPI_EINVAL, "n_packets must be > 0");
    if ((uint64_t)n_packets * fpp > c->desc.max_frames) return fail(DSPI_ERANGE, "%u frames exceed max_frames %u", n_packets * fpp, c->desc.max_frames);
    return DSPI_OK;
}

int dspi_chain_process_device(dspi_chain *c, const void *d_pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t fpp, int32_t *d_spdif,
                              uint32_t *d_pdm, dspi_status *d_status)
{
    int rc = check_process(c, d_pcm, bit_depth, n_packets, fpp);
    if (rc) return rc;
    CU_OK(cudaSetDevice(c->desc.device));
    if (c->desc.arith == DSPI_ARITH_F32_FUSED) return launch_chain<true>(c, d_pcm, bit_depth, n_packets, fpp, d_spdif, d_pdm, d_status);
    return launch_chain<false>(c, d_pcm, bit_depth, n_packets, fpp, d_spdif, d_pdm, d_status);
}

int dspi_chain_process_host(dspi_chain *c, const void *pcm, uint32_t bit_depth, uint32_t n_packets, uint32_t fpp, int32_t *spdif_out,
                            uint32_t *pdm_out, dspi_status *status)
{
    int rc = check_process(c, pcm, bit_depth, n_packets, fpp);
    if (rc) return rc;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t N = c->desc.n_instances, F = (size_t)n_packets * fpp;
    const size_t in_bytes = N * F * (bit_depth == 24 ? 6 : 4), sp_bytes = N * 4 * F * 2 * 4, pd_bytes = N * F * 8 * 4;
    if (in_bytes > c->pcm_bytes) { if (c->d_pcm) cudaFree(c->d_pcm); c->d_pcm = nullptr; c->pcm_bytes = 0; CU_OK(cudaMalloc(&c->d_pcm, in_bytes)); c->pcm_bytes = in_bytes; }
    if (spdif_out && sp_bytes > c->spdif_bytes) { if (c->d_spdif) cudaFree(c->d_spdif); c->d_spdif = nullptr; c->spdif_bytes = 0; CU_OK(cudaMalloc((void **)&c->d_spdif, sp_bytes)); c->spdif_bytes = sp_bytes; }
    if (pdm_out && pd_bytes > c->pdmout_bytes) { if (c->d_pdmout) cudaFree(c->d_pdmout); c->d_pdmout = nullptr; c->pdmout_bytes = 0; CU_OK(cudaMalloc((void **)&c->d_pdmout, pd_bytes)); c->pdmout_bytes = pd_bytes; }
    CU_OK(cudaMemcpyAsync(c->d_pcm, pcm, in_bytes, cudaMemcpyHostToDevice, c->stream));
    rc = dspi_chain_process_device(c, c->d_pcm, bit_depth, n_packets, fpp, spdif_out ? c->d_spdif : nullptr, pdm_out ? c->d_pdmout : nullptr,
                                   status ? c->d_status : nullptr);
    if (rc) return rc;
    if (spdif_out) CU_OK(cudaMemcpyAsync(spdif_out, c->d_spdif, sp_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (pdm_out) CU_OK(cudaMemcpyAsync(pdm_out, c->d_pdmout, pd_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (status) CU_OK(cudaMemcpyAsync(status, c->d_status, N * sizeof(dspi_status), cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chain_sync(dspi_chain *c)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

void *dspi_chain_stream(dspi_chain *c) { return c ? (void *)c->stream : nullptr; }
uint64_t dspi_chain_launch_count(dspi_chain *c) { return c ? c->launches : 0; }

}  // extern "C"
