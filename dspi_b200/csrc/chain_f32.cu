// chain_f32.cu — the whole per-packet float signal chain of one DSPi device, for thousands of
// independent device instances, sm_100a.
//
// Reference: process_audio_packet(), firmware/DSPi/usb_audio.c:500-1317 — float pipeline :560-967,
// single-core branch :874-960; crossfeed.c:132-156; leveller.c:148-262; pdm_generator.c:351-397.
// Chain order (the code's, not the README's): preamp -> loudness -> master EQ -> leveller ->
// crossfeed (+ input peaks) -> matrix -> per-output EQ -> gain x volume -> delay -> peaks ->
// 24-bit words / delta-sigma PDM.
//
// The chain is feed-forward between stages (nothing downstream feeds an upstream stage), so a call is
// run stage by stage over whole slices of packets instead of packet by packet:
//
//   chain_pre_kernel      warp = 16 instances x {L, R}: PCM unpack, preamp, the two loudness shelves; results leave
//                         through a shared-memory transpose as ROWS [2 N][frames] (row = side * N + inst)
//   K1 (eq_f32_kernel.cuh) the 10-band master EQ over those rows — the same TMA-fed packed-FFMA2 kernel
//                         (and run-time specialisation) as the EQ engine, not a second implementation
//   chain_post_kernel     warp = 16 instances x {L, R}: per-packet leveller (stereo-linked, 480-sample
//                         look-ahead ring), input peaks, crossfeed; L <-> R exchange by __shfl_xor(.., 16)
//   chain_mix_kernel      lane = frame: the 2 x 9 matrix, writes output ROWS [9 N][frames]
//   K1                    per-output EQ over the 9 N rows (muted / disabled rows are masked out)
//   chain_outpost_kernel  lane = frame, warp = (instance, packet): gain, delay, peak/clip metering,
//                         float -> 24-bit S/PDIF pairs (coalesced 8-byte stores), Q28 for the modulator.
//                         A delayed sample that lies inside the current call is read from the output
//                         rows; only older history comes from the delay ring in HBM.
//   chain_ring_kernel     once per call: the last <= 4096 post-gain samples go into the rings
//   chain_pdm_kernel      one instance per lane: 256x oversampled 2nd-order delta-sigma (chain_pdm.cuh)
//
// Everything with a serial recurrence but little arithmetic keeps lane = instance; everything without
// one runs lane = frame, fully coalesced; the EQ — 90 % of the arithmetic — runs in the kernel that is
// tuned against the roofline.  All per-instance parameters and states are SoA arrays with the instance
// index innermost.
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

constexpr int kOuts = DSPI_CHAIN_OUTPUTS;
constexpr int kRoles = DSPI_CHAIN_EQ_CHANNELS;
constexpr int kMaxDelay = DSPI_CHAIN_MAX_DELAY;
constexpr int kLa = DSPI_LA_SAMPLES;
constexpr int kPkt = DSPI_PACKET_MAX;
constexpr int kXs = 33;                           // shared-memory column stride: conflict-free for lane = instance AND lane = frame

enum : uint8_t { F_BYPASS_MASTER = 1, F_LOUD = 2, F_XFEED = 4, F_LEV = 8, F_LOOKAHEAD = 16, F_ANY_DELAY = 32, F_SUB_ON = 64 };
enum : uint8_t { O_ENABLED = 1, O_MUTE = 2, O_PAIR_OFF = 4 };

struct ChainDev {
    uint32_t N, N_pad, nb, max_frames, ldF;       // ldF: row stride of mrow / orow / subq (frames, multiple of 4)
    float *preamp;                                // [2][N_pad]
    uint8_t *flags;                               // [N_pad] F_*
    float *loud_c; float *loud_st; uint8_t *loud_byp;   // [2 j][6][N_pad], [2 side][2 j][2][N_pad], [N_pad] bit j
    float *xf;                                    // [7][N_pad] lp_a0 lp_b1 lp_L lp_R ap_a ap_L ap_R
    float *lev_c; float *lev_s; uint32_t *lev_idx; float *lev_la;   // [9][N_pad], [5][N_pad], [N_pad], [2][480][N_pad]
    float *o_gl, *o_gr, *o_gain; uint8_t *o_flags; int32_t *o_dly;  // [9][N_pad]
    float *dline; uint32_t *widx_in, *widx_out;   // [9][N_pad][4096], [N_pad]
    int32_t *pdm;                                 // [9][N_pad] err1 err2 x1 x2 y1 y2 err_acc rng fade_in_pos
    uint16_t *peaks; uint16_t *clip;              // [11][N_pad], [N_pad]
    float *mrow;                                  // [2 N_pad][ldF] master rows, row = side * N_pad + inst
    float *orow;                                  // [9 N_pad][ldF] output rows, row = o * N_pad + inst
    int32_t *subq;                                // [N_pad][ldF] Q28 sub samples for the modulator
    uint8_t *skip_m, *skip_o;                     // [2 N_pad], [9 N_pad]: rows whose EQ is frozen (K1 skip mask)
    // preset-mute envelope (usb_audio.c:456-498): per-instance state and the per-packet volume it produces
    uint32_t *env;                                // [5][N_pad] loading, counter, smooth gain (float bits), sample rate, envelope mode on
    float *vol_base, *vol_master, *o_glin;        // [N_pad] host volume (:569), [N_pad] master volume, [9][N_pad] outputs[o].gain_linear
    float *pmg;                                   // [N_pad] the constant preset_mute_gain of dspi_chain_set_params
    float *vmm;                                   // [packets of the call][N_pad] vol_mul_master (:571) of envelope-mode instances
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
// pre: PCM unpack + preamp (usb_audio.c:591-686), loudness shelves (:689-718) -> master rows
// ---------------------------------------------------------------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(64)
chain_pre_kernel(ChainDev d, const uint8_t *__restrict__ pcm, uint32_t bit_depth, uint32_t f_begin, uint32_t f_end, uint32_t F)
{
    // warp = 16 instances x {L, R}: lane l handles side l >> 4 of instance inst16 + (l & 15).  The two sides of an
    // instance share nothing in this stage (separate shelf states, usb_audio.c:696 / :707), so splitting them over two
    // lanes halves the serial chain per lane and doubles the warps of a stage that is latency-bound on few warps.
    __shared__ float tile_s[2][32][kXs];                  // per warp: [frame][row = side * 16 + instance]
    __shared__ uint32_t pcm_s[2][2][16][49];              // per warp, double-buffered: 16 instances x 32 frames x <= 6 bytes (rows padded to 49 words)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t inst16 = (blockIdx.x * 2 + warp) * 16;
    if (inst16 >= d.N_pad) return;
    const uint32_t side = lane >> 4, li = lane & 15;
    const uint32_t inst = inst16 + li;
    const bool live = inst < d.N;
    const uint32_t Np = d.N_pad;
    float (*tile)[kXs] = tile_s[warp];

    const bool loud_on = d.flags[inst] & F_LOUD;
    const uint8_t loud_byp = d.loud_byp[inst];
    float lc[2][6], ls[2][2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int k = 0; k < 6; k++) lc[j][k] = d.loud_c[(j * 6 + k) * Np + inst];
        ls[j][0] = d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst];
        ls[j][1] = d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst];
    }
    const uint32_t bpf = bit_depth == 24 ? 6u : 4u;
    const float preamp = d.preamp[side * Np + inst];
    const float gain_in = bit_depth == 24 ? __fmul_rn(1.0f / 8388608.0f, preamp)        // usb_audio.c:601-603
                                          : __fmul_rn(1.0f / 32768.0f, preamp);          // :680-681
    // Each instance's packet stream is contiguous ([inst][F] frames of bpf bytes); when every tile starts on
    // a 4-byte boundary the warp fetches the 16 tiles of its instances with coalesced word loads into shared
    // memory and every lane then decodes its own side from there, otherwise lanes read their bytes directly.
    const bool words_ok = ((reinterpret_cast<uintptr_t>(pcm) | ((size_t)F * bpf) | ((size_t)f_begin * bpf)) & 3u) == 0;
    const uint8_t *my_pcm = pcm + (size_t)inst * F * bpf;
    const uint32_t n_inst = min(16u, d.N > inst16 ? d.N - inst16 : 0u);
    // asynchronous fetch of the tile starting at frame f0 into buffer `buf` (one commit group per call).  A last
    // word may run <= 2 bytes past a ragged tile: still inside the PCM buffer, because the very end of the
    // buffer is word-aligned (F * bpf is) and so is every tile start.
    auto fetch = [&](uint32_t f0, int buf) {
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
        fetch(f0 + 32, buf ^ 1);                          // next tile streams in behind this tile's arithmetic
        if (words_ok) {
            cp_async_wait<1>();
            __syncwarp();
            tile_bytes = reinterpret_cast<const uint8_t *>(pcm_s[warp][buf][li]);
        }
        for (uint32_t t = 0; t < nv; t++) {
            const uint8_t *q = tile_bytes + (size_t)t * bpf;
            int32_t sm = 0;
            if (live) {
                if (bit_depth == 24) {
                    const uint8_t *b = q + side * 3;
                    sm = ((int32_t)((uint32_t)b[2] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[0] << 8)) >> 8;
                } else {
                    const uint8_t *b = q + side * 2;
                    sm = (int16_t)((uint16_t)b[0] | (uint16_t)b[1] << 8);
                }
            }
            float v = __fmul_rn((float)sm, gain_in);                         // :645-648 / :683-684
            if (loud_on) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if ((loud_byp >> j) & 1) continue;
                    float &s0 = ls[j][0], &s1 = ls[j][1];
                    const float v3 = __fadd_rn(v, -s1);
                    const float pp = __fmul_rn(lc[j][1], v3);
                    float t2, v1, v2;
                    if (FUSED) {
                        t2 = __fmaf_rn(lc[j][1], s0, s1);
                        v1 = __fmaf_rn(lc[j][0], s0, pp);
                        v2 = __fmaf_rn(lc[j][2], v3, t2);
                    } else {
                        t2 = __fadd_rn(s1, __fmul_rn(lc[j][1], s0));
                        v1 = __fadd_rn(__fmul_rn(lc[j][0], s0), pp);
                        v2 = __fadd_rn(t2, __fmul_rn(lc[j][2], v3));
                    }
                    s0 = __fmaf_rn(2.0f, v1, -s0);
                    s1 = __fmaf_rn(2.0f, v2, -s1);
                    v = fm<FUSED>(lc[j][5], v2, fm<FUSED>(lc[j][3], v, __fmul_rn(lc[j][4], v1)));   // :702
                }
            }
            tile[t][lane] = v;
        }
        __syncwarp();
        // transpose out: lane = frame, one coalesced 128-byte store per (side, instance) row
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
// post: leveller (leveller.c:148-262), input peaks, crossfeed (crossfeed.c:132-156), packet by packet
// ---------------------------------------------------------------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(128)
chain_post_kernel(ChainDev d, uint32_t p0, uint32_t n_packets, uint32_t fpp)
{
    extern __shared__ float smem[];                       // per warp: packet columns [fpp][33] + look-ahead reads [fpp][33]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t side = lane >> 4;
    const uint32_t inst16 = (blockIdx.x * (blockDim.x >> 5) + warp) * 16;
    if (inst16 >= d.N_pad) return;
    const uint32_t inst = inst16 + (lane & 15);
    const uint32_t Np = d.N_pad;
    float *xw = smem + (size_t)warp * 2 * fpp * kXs;      // xw[t * 33 + r]: column r of this warp
    float *xs = xw + lane;                                // own column
    float *hs = xw + (size_t)fpp * kXs + lane;            // held look-ahead samples, own column

    const uint8_t flags = d.flags[inst];
    const bool lev_on = flags & F_LEV, xf_on = flags & F_XFEED, lookahead = flags & F_LOOKAHEAD;
    // crossfeed: this lane owns its side's lowpass / all-pass state
    const float xf_a0 = d.xf[0 * Np + inst], xf_b1 = d.xf[1 * Np + inst], xf_ap = d.xf[4 * Np + inst];
    float xf_lp = d.xf[(2 + side) * Np + inst], xf_as = d.xf[(5 + side) * Np + inst];
    float lvc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) lvc[k] = d.lev_c[k * Np + inst];
    float env = d.lev_s[side * Np + inst];
    float smooth_db = d.lev_s[2 * Np + inst], gain_lin = d.lev_s[3 * Np + inst], gain_prev = d.lev_s[4 * Np + inst];
    uint32_t la_idx = d.lev_idx[inst];
    float *la_buf = d.lev_la + (size_t)side * kLa * Np + inst;

    float peak_in = 0.0f;
    uint16_t clip = 0;
    for (uint32_t p = p0; p < p0 + n_packets; p++) {
        const uint32_t f0 = p * fpp;
        // The look-ahead ring is read one slot per sample, each read just before that slot is overwritten
        // (leveller.c:231-237), and a packet (<= 192 frames) never laps the 480-slot ring: all of this
        // packet's reads are issued now as asynchronous copies, together with the packet itself.
        if (lev_on && lookahead) {
            uint32_t idx = la_idx;
            for (uint32_t i = 0; i < fpp; i++) {
                cp_async_4(hs + i * kXs, la_buf + (size_t)idx * Np);
                if (++idx >= (uint32_t)kLa) idx = 0;
            }
        }
        // packet in: lane = frame, coalesced row reads, transposed into lane-private columns (asynchronous
        // copies again: 32 rows x fpp/32 independent requests in flight instead of one load-store pair at a time)
        for (int r = 0; r < 32; r++) {
            const float *row = d.mrow + ((size_t)(r >> 4) * Np + inst16 + (r & 15)) * d.ldF + f0;
            for (uint32_t t = lane; t < fpp; t += 32) cp_async_4(xw + t * kXs + r, row + t);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();

        // ---- PASS 3 for one sample: input peak, then crossfeed (usb_audio.c:741-749); every lane walks the same shuffle ----
        float pk = 0.0f;
        auto peak_and_crossfeed = [&](uint32_t i, float v) {
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
            xs[i * kXs] = v;
        };

        // ---- PASS 2.5: leveller ----
        if (__any_sync(0xffffffffu, lev_on)) {
            const float a_rms = lvc[0], one_minus = __fadd_rn(1.0f, -a_rms);
            float e = env;
            for (uint32_t i = 0; i < fpp; i++) {                             // leveller.c:161-166
                const float s = xs[i * kXs];
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
            // the leveller's per-sample part and PASS 3 share one loop: ramp, look-ahead exchange and peak limit of sample
            // i+1 do not depend on the crossfeed recurrence of sample i, so the serial chains overlap
            for (uint32_t i = 0; i < fpp; i++) {                             // :228-259, then usb_audio.c:741-749
                const float x0 = xs[i * kXs];
                float o = x0;
                if (lev_on && lookahead) {
                    const float held = hs[i * kXs];
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
                gain = __fadd_rn(gain, gain_step);
                peak_and_crossfeed(i, lev_on ? __fmul_rn(o, g) : x0);
            }
            if (lev_on) {
                env = e;
                smooth_db = new_smooth;
                gain_prev = gain_lin;
                gain_lin = new_gain;
            }
        } else {
            for (uint32_t i = 0; i < fpp; i++) peak_and_crossfeed(i, xs[i * kXs]);
        }
        peak_in = pk;                                                        // peaks describe the last packet
        if (pk > 1.001f) clip |= (uint16_t)(1u << side);                     // config.h:53
        __syncwarp();
        // packet out: back into the same rows, coalesced
        for (uint32_t t = lane; t < fpp; t += 32) {
#pragma unroll 8
            for (int r = 0; r < 32; r++)
                d.mrow[((size_t)(r >> 4) * Np + inst16 + (r & 15)) * d.ldF + f0 + t] = xw[t * kXs + r];
        }
        __syncwarp();
    }

    // ---- state back ----
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
    // clip_flags bits 0/1: the two sides of one instance sit in lanes l and l+16; two instances share a 32-bit word
    const uint16_t clip_other = (uint16_t)__shfl_xor_sync(0xffffffffu, (uint32_t)clip, 16);
    if (side == 0 && (clip | clip_other)) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), (unsigned int)(clip | clip_other) << (16 * (inst & 1)));
}

// ---------------------------------------------------------------------------------------------
// matrix mix (usb_audio.c:753-779): lane = frame
// ---------------------------------------------------------------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(256)
chain_mix_kernel(ChainDev d, uint32_t f_begin, uint32_t f_end)
{
    const int lane = threadIdx.x & 31;
    constexpr int kB = 4;                                  // frames per lane per unit: kB independent loads in flight
    const uint32_t n_tiles = (f_end - f_begin + 32 * kB - 1) / (32 * kB);
    const uint64_t units = (uint64_t)d.N * n_tiles;
    const uint32_t Np = d.N_pad;
    for (uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t inst = (uint32_t)(u / n_tiles), tile = (uint32_t)(u % n_tiles);
        const uint32_t fbase = f_begin + tile * 32 * kB + lane;
        float l[kB], r[kB];
#pragma unroll
        for (int j = 0; j < kB; j++) {
            const uint32_t f = fbase + 32 * j;
            l[j] = f < f_end ? d.mrow[(size_t)inst * d.ldF + f] : 0.0f;
            r[j] = f < f_end ? d.mrow[((size_t)Np + inst) * d.ldF + f] : 0.0f;
        }
#pragma unroll
        for (int o = 0; o < kOuts; o++) {
            const bool enabled = d.o_flags[o * Np + inst] & O_ENABLED;
            const float gl = d.o_gl[o * Np + inst], gr = d.o_gr[o * Np + inst];
            const int mixcase = !enabled ? 0 : (gl != 0.0f && gr != 0.0f) ? 3 : (gl != 0.0f) ? 1 : (gr != 0.0f) ? 2 : 0;   // :767-778
#pragma unroll
            for (int j = 0; j < kB; j++) {
                float v;
                if (mixcase == 3) v = fm<FUSED>(l[j], gl, __fmul_rn(r[j], gr));          // :769
                else if (mixcase == 1) v = __fmul_rn(l[j], gl);
                else if (mixcase == 2) v = __fmul_rn(r[j], gr);
                else v = 0.0f;
                const uint32_t f = fbase + 32 * j;
                if (f < f_end) d.orow[((size_t)o * Np + inst) * d.ldF + f] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// outputs after the EQ: gain (:885-894), delay (:898-912), peaks (:914-923), 24-bit / Q28 (:925-959)
// ---------------------------------------------------------------------------------------------
// update_preset_mute_envelope() (usb_audio.c:466-498) for every packet of the call, one instance per thread, and the
// volume chain of :569-571 that depends on it: vmm[p] = (vol_base * g_p) * master_volume_linear.  Same operations, same
// order as dspi_preset_mute_step() on the host.
__global__ void chain_env_kernel(ChainDev d, uint32_t n_packets, uint32_t fpp)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t Np = d.N_pad;
    if (inst >= d.N || !d.env[4 * Np + inst]) return;
    uint32_t loading = d.env[0 * Np + inst], counter = d.env[1 * Np + inst];
    float g = __uint_as_float(d.env[2 * Np + inst]);
    const uint32_t fs = d.env[3 * Np + inst];
    unsigned long long ts = ((unsigned long long)fs * 8ull + 999ull) / 1000ull;          // :459-464
    if (ts < 1ull) ts = 1ull;
    if (ts > 0xFFFFFFFFull) ts = 0xFFFFFFFFull;
    float step = __fdiv_rn((float)fpp, (float)(uint32_t)ts);                             // :486
    if (step > 1.0f) step = 1.0f;
    const float vol_base = d.vol_base[inst], master = d.vol_master[inst];
    for (uint32_t p = 0; p < n_packets; p++) {
        const bool active = loading != 0;                                                // :469
        if (active) {
            if (counter > fpp) counter -= fpp;
            else { counter = 0; loading = 0; }
        }
        const float target = active ? 0.0f : 1.0f;
        if (g < target)      { g = __fadd_rn(g, step);  if (g > target) g = target; }
        else if (g > target) { g = __fadd_rn(g, -step); if (g < target) g = target; }
        d.vmm[(size_t)p * Np + inst] = __fmul_rn(__fmul_rn(vol_base, g), master);        // :570-571
    }
    d.env[0 * Np + inst] = loading;
    d.env[1 * Np + inst] = counter;
    d.env[2 * Np + inst] = __float_as_uint(g);
}

// Mass reconfiguration of the dynamics stages on the device (SURVEY f-1): what the main loop does for one instance when
// crossfeed_update_pending / leveller_update_pending / loudness_recompute_pending are set (main.c:868-895) plus
// audio_set_volume() (usb_audio.c:428-440), one instance per thread, written straight into the engine's arrays.
__global__ void chain_dynamics_kernel(ChainDev d, uint32_t inst0, uint32_t n, const dspi_dynamics_config *__restrict__ cfgs, float fs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t inst = inst0 + i, Np = d.N_pad;
    const dspi_dynamics_config cfg = cfgs[i];
    uint8_t flags = d.flags[inst] & (uint8_t)~(F_XFEED | F_LEV | F_LOOKAHEAD | F_LOUD);
    // crossfeed_compute_coefficients(): coefficients and CLEARED filter state (crossfeed.c:110-126)
    float a0, b1, ap;
    dyn::crossfeed_coeffs(cfg.crossfeed, fs, a0, b1, ap);
    d.xf[0 * Np + inst] = a0; d.xf[1 * Np + inst] = b1; d.xf[4 * Np + inst] = ap;
    d.xf[2 * Np + inst] = 0.0f; d.xf[3 * Np + inst] = 0.0f; d.xf[5 * Np + inst] = 0.0f; d.xf[6 * Np + inst] = 0.0f;
    if (cfg.crossfeed.enabled) flags |= F_XFEED;                             // crossfeed_bypassed = !enabled, main.c:882
    // leveller_compute_coefficients(); leveller_bypassed = !enabled, main.c:886-894
    float lv[9];
    dyn::leveller_coeffs(cfg.leveller, fs, lv);
#pragma unroll
    for (int k = 0; k < 9; k++) d.lev_c[k * Np + inst] = lv[k];
    if (cfg.leveller.enabled) flags |= F_LEV;
    if (cfg.leveller.lookahead) flags |= F_LOOKAHEAD;
    // audio_set_volume(): vol_mul and the loudness row of that volume step; loudness_recompute_table() for that row
    uint32_t row;
    const int16_t vol_mul = dyn::host_volume(cfg.volume_8_8, row);
    float lo_db, hi_db;
    dyn::loudness_row_gains((int)row, cfg.loudness_ref_spl, cfg.loudness_intensity_pct, lo_db, hi_db);
    float c[6];
    bool byp;
    uint8_t lb = 0;
    const float lfs = fs < 1.0f ? 48000.0f : fs;                             // loudness.c:171
    dyn::shelf_svf(200.0f, 0.707f, lo_db, false, lfs, c, byp);
    if (byp) lb |= 1;
#pragma unroll
    for (int k = 0; k < 6; k++) d.loud_c[(0 * 6 + k) * Np + inst] = c[k];
    dyn::shelf_svf(6000.0f, 0.707f, hi_db, true, lfs, c, byp);
    if (byp) lb |= 2;
#pragma unroll
    for (int k = 0; k < 6; k++) d.loud_c[(1 * 6 + k) * Np + inst] = c[k];
    d.loud_byp[inst] = lb;
    if (cfg.loudness_enabled) flags |= F_LOUD;
    d.flags[inst] = flags;
    // host volume -> output gains (usb_audio.c:569-571, 886-887)
    const float vol_base = cfg.host_mute ? 0.0f : __fmul_rn((float)vol_mul, 1.0f / 32768.0f);
    d.vol_base[inst] = vol_base;
    const float vmm = __fmul_rn(__fmul_rn(vol_base, d.pmg[inst]), d.vol_master[inst]);
    for (int o = 0; o < kOuts; o++)
        d.o_gain[o * Np + inst] = (d.o_flags[o * Np + inst] & O_MUTE) ? 0.0f : __fmul_rn(d.o_glin[o * Np + inst], vmm);
}

// post-gain sample of output row `o` (what the delay line stores)
__device__ __forceinline__ float out_gain(float v, bool enabled, float gain)
{
    if (enabled) {
        if (gain == 0.0f) v = 0.0f;
        else if (gain != 1.0f) v = __fmul_rn(v, gain);
    }
    return v;
}

struct OutCfg {
    bool enabled, pair_off, delay_on, mute;
    float gain;                        // constant gain of the call (no envelope)
    float glin;                        // outputs[o].gain_linear
    const float *vmm;                  // envelope mode: vol_mul_master per packet, stride N_pad; else nullptr
    uint32_t fpp, Np;
    uint32_t dl;                       // delay & (MAX - 1): MAX aliases to 0 (SURVEY a-10)
    const float *row;                  // orow row of this (output, instance)
    const float *ring;
};

__device__ __forceinline__ OutCfg out_cfg(const ChainDev &d, uint32_t o, uint32_t inst, bool any_delay, uint32_t fpp)
{
    OutCfg c;
    const uint32_t Np = d.N_pad;
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
    c.delay_on = any_delay && dly > 0;                                       // usb_audio.c:898-901
    c.dl = (uint32_t)dly & (kMaxDelay - 1);
    c.row = d.orow + ((size_t)o * Np + inst) * d.ldF;
    c.ring = d.dline + ((size_t)o * Np + inst) * kMaxDelay;
    return c;
}

// output gain in force at frame T of the call (usb_audio.c:886-887): constant, or following the envelope packet by packet
__device__ __forceinline__ float gain_at(const OutCfg &c, uint32_t T)
{
    if (!c.vmm) return c.gain;
    return c.mute ? 0.0f : __fmul_rn(c.glin, c.vmm[(size_t)(T / c.fpp) * c.Np]);
}

// the sample output `c` emits at frame T of this call (T counted from the start of the call):
// write-then-read per sample (:902-909) means frame T emits the post-gain sample of frame T - dl;
// inside the call that sample is still in the output rows, before it only the ring has it
__device__ __forceinline__ float out_sample(const OutCfg &c, uint32_t T, uint32_t widx0)
{
    if (!c.delay_on) return out_gain(c.row[T], c.enabled, gain_at(c, T));
    if (T >= c.dl) return out_gain(c.row[T - c.dl], c.enabled, gain_at(c, T - c.dl));
    return c.ring[(widx0 + T - c.dl) & (kMaxDelay - 1)];
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, s);
        if (o > v) v = o;
    }
    return v;
}

__global__ void __launch_bounds__(256)
chain_outpost_kernel(ChainDev d, uint32_t p0, uint32_t n_packets, uint32_t fpp, uint32_t F, int32_t *__restrict__ spdif_out)
{
    const int lane = threadIdx.x & 31;
    const uint64_t units = (uint64_t)d.N * n_packets;
    const uint32_t Np = d.N_pad;
    for (uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t inst = (uint32_t)(u / n_packets), p = p0 + (uint32_t)(u % n_packets);
        const uint32_t f0 = p * fpp;
        const bool last = p == p0 + n_packets - 1;
        const bool any_delay = d.flags[inst] & F_ANY_DELAY;
        const uint32_t widx0 = d.widx_in[inst];
        unsigned int clip = 0;
        for (int k = 0; k <= 4; k++) {                                        // four S/PDIF pairs, then the sub alone
            const bool is_sub = k == 4;
            const uint32_t oa = 2 * k, ob = is_sub ? oa : oa + 1;
            const OutCfg ca = out_cfg(d, oa, inst, any_delay, fpp), cb = out_cfg(d, ob, inst, any_delay, fpp);
            float pka = 0.0f, pkb = 0.0f;
            constexpr int kB = 4;                                            // independent loads in flight per lane
            for (uint32_t tb = lane; tb < fpp; tb += 32 * kB) {
                float xa[kB], xb[kB];
#pragma unroll
                for (int j = 0; j < kB; j++) {
                    const uint32_t t = tb + 32 * j;
                    xa[j] = t < fpp ? out_sample(ca, f0 + t, widx0) : 0.0f;
                    xb[j] = (!is_sub && t < fpp) ? out_sample(cb, f0 + t, widx0) : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < kB; j++) {
                    const uint32_t t = tb + 32 * j, T = f0 + t;
                    if (t >= fpp) break;
                    const float aa = fabsf(xa[j]), ab = fabsf(xb[j]);
                    if (aa > pka) pka = aa;
                    if (ab > pkb) pkb = ab;
                    if (is_sub) {
                        if (ca.enabled) d.subq[(size_t)inst * d.ldF + T] = __float2int_rz(__fmul_rn(xa[j], 268435456.0f));   // :953 (saturating)
                    } else if (spdif_out) {
                        int2 w = make_int2(0, 0);
                        if (!ca.pair_off) {                                  // :930-939 (pair_off is a property of the pair)
                            w.x = __float2int_rz(__fmul_rn(fmaxf(-1.0f, fminf(1.0f, xa[j])), 8388607.0f));
                            w.y = __float2int_rz(__fmul_rn(fmaxf(-1.0f, fminf(1.0f, xb[j])), 8388607.0f));
                        }
                        *reinterpret_cast<int2 *>(spdif_out + (((size_t)inst * 4 + k) * F + T) * 2) = w;
                    }
                }
            }
            pka = warp_max(pka);
            pkb = warp_max(pkb);
            if (lane == 0) {
                if (last) {
                    uint16_t pq = (uint16_t)__fmul_rn(fminf(1.0f, pka), 32767.0f);              // :921 / :950
                    if (is_sub && !ca.enabled) pq = 0;                                           // :957
                    d.peaks[(2 + oa) * Np + inst] = pq;
                    if (!is_sub) d.peaks[(2 + ob) * Np + inst] = (uint16_t)__fmul_rn(fminf(1.0f, pkb), 32767.0f);
                }
                if (pka > 1.001f && (!is_sub || ca.enabled)) clip |= 1u << (2 + oa);
                if (!is_sub && pkb > 1.001f) clip |= 1u << (2 + ob);
            }
        }
        if (lane == 0 && clip) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), clip << (16 * (inst & 1)));
    }
}

// once per call, after every outpost launch of the call: the delay rings take the last <= 4096 post-gain
// samples (older writes of this call would have been overwritten anyway), the shared write index advances
__global__ void __launch_bounds__(256)
chain_ring_kernel(ChainDev d, uint32_t F, uint32_t fpp)
{
    const int lane = threadIdx.x & 31;
    const uint64_t units = (uint64_t)d.N * kOuts;
    const uint32_t Np = d.N_pad;
    for (uint64_t u = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint32_t inst = (uint32_t)(u / kOuts), o = (uint32_t)(u % kOuts);
        const bool any_delay = d.flags[inst] & F_ANY_DELAY;
        const uint32_t widx0 = d.widx_in[inst];
        const OutCfg c = out_cfg(d, o, inst, any_delay, fpp);
        if (c.delay_on) {                                                    // outputs without delay never touch their ring
            float *ring = d.dline + ((size_t)o * Np + inst) * kMaxDelay;
            for (uint32_t T = (F > (uint32_t)kMaxDelay ? F - kMaxDelay : 0u) + lane; T < F; T += 32)
                ring[(widx0 + T) & (kMaxDelay - 1)] = out_gain(c.row[T], c.enabled, gain_at(c, T));
        }
        if (o == 0 && lane == 0) d.widx_out[inst] = any_delay ? (widx0 + F) & (kMaxDelay - 1) : widx0;   // :911, once per packet
    }
}

// ---------------------------------------------------------------------------------------------
// delta-sigma PDM (chain_pdm.cuh): one instance per lane
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
chain_pdm_kernel(ChainDev d, uint32_t f_begin, uint32_t f_end, uint32_t F, uint32_t *__restrict__ pdm_out)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= d.N) return;
    if (!(d.flags[inst] & F_SUB_ON)) return;                                 // usb_audio.c:944
    pdm_modulate_frames(d.pdm, d.subq + (size_t)inst * d.ldF, 1, d.N_pad, inst, f_begin, f_end, F, pdm_out);
}

// filters[][] of n instances (instance-major AoS) <-> the mirrors of the two EQ engines (channel = role' * N_pad + inst)
__global__ void chain_scatter_kernel(const dspi_biquad_f32 *__restrict__ aos, uint32_t inst0, uint32_t n, uint32_t Np, dspi_biquad_f32 *__restrict__ m_aos,
                                     dspi_biquad_f32 *__restrict__ o_aos, int to_mirrors)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kRoles * kMaxBands) return;
    const uint32_t b = i % kMaxBands, role = (i / kMaxBands) % kRoles, inst = inst0 + i / (kMaxBands * kRoles);
    dspi_biquad_f32 *chain_q = const_cast<dspi_biquad_f32 *>(aos) + ((size_t)inst * kRoles + role) * kMaxBands + b;
    dspi_biquad_f32 *eng_q = role < 2 ? m_aos + ((size_t)role * Np + inst) * kMaxBands + b : o_aos + ((size_t)(role - 2) * Np + inst) * kMaxBands + b;
    if (to_mirrors) *eng_q = *chain_q;
    else *chain_q = *eng_q;
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
    dspi_eq *eq_m, *eq_o;            // K1 engines over the master rows (2 N_pad channels) and the output rows (9 N_pad)
    std::vector<void *> allocs;
    uint64_t launches;
    void *d_pcm; size_t pcm_bytes;   // host-path staging
    int32_t *d_spdif; size_t spdif_bytes;
    uint32_t *d_pdmout; size_t pdmout_bytes;
    dspi_status *d_status;
    uint32_t env_instances;          // instances in envelope mode (0: the envelope kernel and its table are not needed)
    uint32_t vmm_packets;            // capacity of d.vmm in packets
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
    auto post = dspi::chain_post_kernel<FUSED>;
    const size_t post_smem = (size_t)4 * 2 * fpp * dspi::kXs * 4;           // 4 warps x (packet + look-ahead columns)
    static dspi::PerDeviceOnce once;                                // per instantiation (flavour)
    int dev = 0;
    if (once.needs(&dev)) {
        CU_OK(cudaFuncSetAttribute(post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)4 * 2 * dspi::kPkt * dspi::kXs * 4)));
        once.mark(dev);
    }
    // Stage pipeline over packet slices on three streams (chain_streams.cuh): front stages of slice
    // i+1 overlap the output stages of slice i and the modulator of slice i-1.
    dspi::ChainStreams &st = c->st;
    uint32_t slice_bounds[dspi::ChainStreams::kMaxSlices + 1];
    const uint32_t n_slices = (uint32_t)dspi::ChainStreams::plan_slices(n_packets, slice_bounds);
    if (c->env_instances) {                                                  // preset-mute envelope: this call's per-packet volumes
        if (c->vmm_packets < n_packets) {
            CU_OK(cudaStreamSynchronize(c->stream));
            if (c->d.vmm) CU_OK(cudaFree(c->d.vmm));
            c->d.vmm = nullptr; c->vmm_packets = 0;
            CU_OK(cudaMalloc((void **)&c->d.vmm, (size_t)n_packets * c->d.N_pad * sizeof(float)));
            c->vmm_packets = n_packets;
        }
        dspi::chain_env_kernel<<<(c->d.N + 127) / 128, 128, 0, c->stream>>>(c->d, n_packets, fpp);
        CU_OK(cudaGetLastError());
        c->launches++;
    }
    const ChainDev d = c->d;
    const uint32_t n_sms = st.rest_sms ? st.rest_sms : 148;     // SMs the streaming stages run on (chain_streams.cuh)
    static const uint32_t kStreamCtas = [] { const char *e = getenv("DSPI_CHAIN_CTAS"); const int v = e ? atoi(e) : 0; return (uint32_t)(v >= 1 && v <= 8 ? v : 8); }();   // streaming CTAs (256 threads) per SM
    CU_OK(cudaEventRecord(st.ev_begin, c->stream));
    CU_OK(cudaStreamWaitEvent(st.s_front, st.ev_begin, 0));
    for (uint32_t sl = 0; sl < n_slices; sl++) {
        const uint32_t p0 = slice_bounds[sl], p1 = slice_bounds[sl + 1];
        const uint32_t fb = p0 * fpp, fe = p1 * fpp;
        int rc;
        // ---- front: unpack + loudness -> master EQ (K1) -> leveller + crossfeed
        dspi::chain_pre_kernel<FUSED><<<(d.N_pad / 16 + 1) / 2, 64, 0, st.s_front>>>(d, (const uint8_t *)d_pcm, bit_depth, fb, fe, F);
        CU_OK(cudaGetLastError());
        if ((rc = dspi::eq_process_on(c->eq_m, d.mrow + fb, fe - fb, d.ldF, st.s_front)) != DSPI_OK) return rc;
        post<<<(d.N_pad / 16 + 3) / 4, 128, post_smem, st.s_front>>>(d, p0, p1 - p0, fpp);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(st.ev_front[sl], st.s_front));
        // ---- outputs: matrix -> per-output EQ (K1) -> gain / delay / metering / conversion
        CU_OK(cudaStreamWaitEvent(st.s_out, st.ev_front[sl], 0));
        dspi::chain_mix_kernel<FUSED><<<n_sms * kStreamCtas, 256, 0, st.s_out>>>(d, fb, fe);
        CU_OK(cudaGetLastError());
        if ((rc = dspi::eq_process_on(c->eq_o, d.orow + fb, fe - fb, d.ldF, st.s_out)) != DSPI_OK) return rc;
        dspi::chain_outpost_kernel<<<n_sms * kStreamCtas, 256, 0, st.s_out>>>(d, p0, p1 - p0, fpp, F, d_spdif);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(st.ev_out[sl], st.s_out));
        // ---- modulator
        CU_OK(cudaStreamWaitEvent(st.s_pdm, st.ev_out[sl], 0));
        dspi::chain_pdm_kernel<<<(d.N + 127) / 128, 128, 0, st.s_pdm>>>(d, fb, fe, F, d_pdm);
        CU_OK(cudaGetLastError());
        c->launches += 5;
    }
    dspi::chain_ring_kernel<<<n_sms * kStreamCtas, 256, 0, st.s_out>>>(d, F, fpp);     // after the last outpost launch (stream order)
    CU_OK(cudaGetLastError());
    c->launches++;
    std::swap(c->d.widx_in, c->d.widx_out);
    CU_OK(cudaEventRecord(st.ev_aux, st.s_out));                             // ring update done
    CU_OK(cudaStreamWaitEvent(c->stream, st.ev_aux, 0));
    // the last modulator launch is ordered after every other stage launch of this call
    CU_OK(cudaEventRecord(st.ev_done, st.s_pdm));
    CU_OK(cudaStreamWaitEvent(c->stream, st.ev_done, 0));                    // later work on the engine stream sees all outputs
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
    c->eq_m = c->eq_o = nullptr;
    c->d_aos = nullptr; c->launches = 0; c->d_pcm = nullptr; c->pcm_bytes = 0; c->d_spdif = nullptr; c->spdif_bytes = 0;
    c->d_pdmout = nullptr; c->pdmout_bytes = 0; c->d_status = nullptr;
    c->env_instances = 0; c->vmm_packets = 0;
    c->desc = *desc;
    ChainDev &d = c->d;
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
        ed.arith = desc->arith; ed.n_bands = desc->n_bands; ed.device = desc->device;
        ed.n_channels = 2 * d.N_pad;
        int rc = dspi_eq_create(&c->eq_m, &ed);
        ed.n_channels = dspi::kOuts * d.N_pad;
        if (rc == DSPI_OK) rc = dspi_eq_create(&c->eq_o, &ed);
        if (rc != DSPI_OK) { dspi_chain_destroy(c); return rc; }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = c->st.create(desc->device, desc->n_instances);
#define TRY(x) if (e == cudaSuccess) e = (x)
    TRY(dev_alloc(c, &c->d_aos, Np * dspi::kRoles * DSPI_MAX_BANDS));
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
    std::vector<float> preamp(2 * n), loud_c(12 * n), xf(7 * n), lev_c(9 * n), gl(9 * n), gr(9 * n), gain(9 * n), glin(9 * n), vbase(n), vmaster(n), pmgv(n);
    std::vector<uint8_t> flags(n), loud_byp(n), oflags(9 * n), skip_m(2 * n), skip_o(9 * n);
    std::vector<int32_t> dly(9 * n);
    std::vector<float> xf_cur(7 * n);
    CU_OK(cudaMemcpy2DAsync(xf_cur.data(), (size_t)n * 4, d.xf + inst0, Np * 4, (size_t)n * 4, 7, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; i++) {
        const dspi_chain_params_f32 &p = params[i];
        // usb_audio.c:569-571
        float vol_mul = p.host_mute ? 0.0f : (float)p.host_vol_mul * (1.0f / 32768.0f);
        vbase[i] = vol_mul;
        vmaster[i] = p.master_volume_linear;
        pmgv[i] = p.preset_mute_gain;
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
            glin[o * n + i] = oc.gain_linear;
            uint8_t f = (oc.enabled ? dspi::O_ENABLED : 0) | (oc.mute ? dspi::O_MUTE : 0);
            if (o < dspi::kOuts - 1) {
                const int partner = o ^ 1;
                if (!oc.enabled && !p.matrix.outputs[partner].enabled) f |= dspi::O_PAIR_OFF;    // :930-933
            }
            oflags[o * n + i] = f;
            skip_o[o * n + i] = (!oc.enabled || oc.mute) ? 1 : 0;            // :878-884: state frozen
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
        skip_m[0 * n + i] = skip_m[1 * n + i] = p.bypass_master_eq ? 1 : 0;   // :721-728
        loud_byp[i] = (p.loudness[0].bypass ? 1 : 0) | (p.loudness[1].bypass ? 2 : 0);
        for (int j = 0; j < 2; j++) {
            const float v[6] = { p.loudness[j].sva1, p.loudness[j].sva2, p.loudness[j].sva3, p.loudness[j].svm0, p.loudness[j].svm1, p.loudness[j].svm2 };
            for (int k = 0; k < 6; k++) loud_c[(j * 6 + k) * n + i] = v[k];
        }
        const float xv[7] = { p.crossfeed.lp_a0, p.crossfeed.lp_b1, p.crossfeed.lp_state_L, p.crossfeed.lp_state_R,
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
    CU_OK(put(d.loud_c, loud_c.data(), 12, 4));
    CU_OK(put(d.loud_byp, loud_byp.data(), 1, 1));
    CU_OK(put(d.xf, xf.data(), 7, 4));
    CU_OK(put(d.lev_c, lev_c.data(), 9, 4));
    CU_OK(put(d.o_gl, gl.data(), 9, 4));
    CU_OK(put(d.o_gr, gr.data(), 9, 4));
    CU_OK(put(d.o_gain, gain.data(), 9, 4));
    CU_OK(put(d.o_glin, glin.data(), 9, 4));
    CU_OK(put(d.vol_base, vbase.data(), 1, 4));
    CU_OK(put(d.vol_master, vmaster.data(), 1, 4));
    CU_OK(put(d.pmg, pmgv.data(), 1, 4));
    CU_OK(put(d.o_flags, oflags.data(), 9, 1));
    CU_OK(put(d.o_dly, dly.data(), 9, 4));
    CU_OK(put(d.skip_m, skip_m.data(), 2, 1));
    CU_OK(put(d.skip_o, skip_o.data(), 9, 1));
    CU_OK(cudaStreamSynchronize(c->stream));
    int rc = dspi::eq_set_skip(c->eq_m, d.skip_m, c->stream);
    if (rc == DSPI_OK) rc = dspi::eq_set_skip(c->eq_o, d.skip_o, c->stream);
    return rc;
}

/* preset-mute envelope of instances [inst0, inst0+n): states == NULL leaves envelope mode (the constant
 * preset_mute_gain of dspi_chain_set_params applies again) */
int dspi_chain_set_preset_mute(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_preset_mute *states, uint32_t sample_rate_hz)
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

int dspi_chain_get_preset_mute(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_preset_mute *states)
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
int dspi_chain_set_dynamics_device(dspi_chain *c, uint32_t inst0, uint32_t n, const dspi_dynamics_config *cfgs, float sample_rate)
{
    if (!c || !cfgs) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    dspi_dynamics_config *d_cfg = nullptr;
    CU_OK(cudaMalloc((void **)&d_cfg, (size_t)n * sizeof(*cfgs)));
    cudaError_t e = cudaMemcpyAsync(d_cfg, cfgs, (size_t)n * sizeof(*cfgs), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        dspi::chain_dynamics_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->d, inst0, n, d_cfg, sample_rate);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_cfg);
    if (e != cudaSuccess) return fail(DSPI_ECUDA, "dynamics coefficient generation: %s", cudaGetErrorString(e));
    c->launches++;
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
    const uint32_t Np = c->d.N_pad, items = n * dspi::kRoles * DSPI_MAX_BANDS;
    dspi::chain_scatter_kernel<<<(items + 255) / 256, 256, 0, c->stream>>>(c->d_aos, inst0, n, Np, (dspi_biquad_f32 *)dspi::eq_aos_mirror(c->eq_m),
                                                                          (dspi_biquad_f32 *)dspi::eq_aos_mirror(c->eq_o), 1);
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

int dspi_chain_set_eq_params_device(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_eq_param *recipes, float sample_rate)
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

int dspi_chain_download_biquads(dspi_chain *c, uint32_t inst0, uint32_t n, dspi_biquad_f32 *biquads)
{
    if (!c || !biquads) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const uint32_t Np = c->d.N_pad, items = n * dspi::kRoles * DSPI_MAX_BANDS;
    for (int role = 0; role < dspi::kRoles; role++) {
        int rc = role < 2 ? dspi::eq_unpack_range(c->eq_m, role * Np + inst0, n, c->stream)
                          : dspi::eq_unpack_range(c->eq_o, (role - 2) * Np + inst0, n, c->stream);
        if (rc) return rc;
    }
    dspi::chain_scatter_kernel<<<(items + 255) / 256, 256, 0, c->stream>>>(c->d_aos, inst0, n, Np, (dspi_biquad_f32 *)dspi::eq_aos_mirror(c->eq_m),
                                                                          (dspi_biquad_f32 *)dspi::eq_aos_mirror(c->eq_o), 0);
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
    if (pdm_out && pd_bytes > c->pdmout_bytes) { if (c->d_pdmout) cudaFree(c->d_pdmout); c->d_pdmout = nullptr; c->pdmout_bytes = 0; CU_OK(cudaMalloc((void **)&c->d_pdmout, pd_bytes)); c->pdmout_bytes = pd_bytes; CU_OK(cudaMemsetAsync(c->d_pdmout, 0, pd_bytes, c->stream)); }
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


// ---- checkpoint / resume: everything a later process call depends on besides the parameters ----------------
static void state_sections(dspi_chain *c, std::vector<std::pair<void *, size_t>> &v)
{
    const size_t Np = c->d.N_pad;
    v.push_back({ c->d.loud_st, 8 * Np * 4 });
    v.push_back({ c->d.xf, 7 * Np * 4 });                                  // crossfeed coefficients and state (CrossfeedState)
    v.push_back({ c->d.lev_s, 5 * Np * 4 });
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

size_t dspi_chain_state_size(dspi_chain *c)
{
    if (!c) return 0;
    std::vector<std::pair<void *, size_t>> v;
    state_sections(c, v);
    size_t n = sizeof(StateHeader);
    for (auto &s : v) n += s.second;
    return n;
}

int dspi_chain_state_export(dspi_chain *c, void *blob, size_t cap)
{
    if (!c || !blob) return fail(DSPI_EINVAL, "null argument");
    const size_t need = dspi_chain_state_size(c);
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

int dspi_chain_state_import(dspi_chain *c, const void *blob, size_t len)
{
    if (!c || !blob) return fail(DSPI_EINVAL, "null argument");
    std::vector<std::pair<void *, size_t>> v;
    state_sections(c, v);
    StateHeader h;
    if (len < sizeof(h)) return fail(DSPI_EINVAL, "state blob too short");
    memcpy(&h, blob, sizeof(h));
    if (h.magic != kStateMagic || h.version != 1u) return fail(DSPI_EINVAL, "not a dspi_b200 state blob (magic %08x version %u)", h.magic, h.version);
    if (h.arith != c->desc.arith || h.n_instances != c->desc.n_instances || h.n_bands != c->desc.n_bands || h.n_sections != v.size() ||
        h.bytes != dspi_chain_state_size(c) || len < h.bytes)
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

int dspi_chain_sync(dspi_chain *c)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

void *dspi_chain_stream(dspi_chain *c) { return c ? (void *)c->stream : nullptr; }
/* SMs reserved for the modulator / left to every other stage (0, 0: no partition, see chain_streams.cuh) */
int dspi_chain_sm_partition(dspi_chain *c, uint32_t *pdm_sms, uint32_t *rest_sms)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    if (pdm_sms) *pdm_sms = c->st.pdm_sms;
    if (rest_sms) *rest_sms = c->st.rest_sms;
    return DSPI_OK;
}

uint64_t dspi_chain_launch_count(dspi_chain *c)
{
    return c ? c->launches + dspi_eq_launch_count(c->eq_m) + dspi_eq_launch_count(c->eq_o) : 0;
}

}  // extern "C"
