// chain_q28.cu — the whole per-packet signal chain in the RP2040's Q28 fixed-point arithmetic, for
// thousands of independent device instances (2 inputs -> 5 outputs each), bit-exact, sm_100a.
//
// Reference: process_audio_packet(), firmware/DSPi/usb_audio.c:968-1283 (single-core branch
// :1191-1276); fast_mul_q28 dsp_pipeline.c:47-58; fast_mul_q15 config.h:556-567; cascade
// dsp_process_rp2040.S:225-394; crossfeed.c:161-180; leveller.c:275-389; PDM pdm_generator.c:351-397.
//
// Same decomposition as chain_f32.cu (front: 16 instances x {L,R} per warp with __shfl_xor(..,16)
// for the stereo-linked leveller and the crossfeed mix; outputs: one output index x 32 instances per
// warp; modulator: one instance per lane).  The EQ runs in the reference's own loop order — band
// outer, the packet's samples inner — over a lane-private shared-memory column, loading the five
// coefficients and two state words of a band once per packet from an SoA store in HBM; everything is
// integer-pipe bound (≈ 27 integer ops per band-sample), so registers are kept low for occupancy.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "eq_kernels.cuh"
#include "chain_pdm.cuh"
#include "chain_streams.cuh"

namespace dspi {
namespace {

constexpr int kOuts = DSPI_CHAINQ_OUTPUTS;
constexpr int kRoles = DSPI_CHAINQ_EQ_CHANNELS;
constexpr int kMaxDelay = DSPI_CHAINQ_MAX_DELAY;
constexpr int kLa = DSPI_LA_SAMPLES;
constexpr int kPkt = DSPI_PACKET_MAX;
constexpr int32_t kUnity = 1 << 28;
constexpr int32_t kClipThresh = (1 << 28) + 268;              // config.h:54

enum : uint8_t { F_BYPASS_MASTER = 1, F_LOUD = 2, F_XFEED = 4, F_LEV = 8, F_LOOKAHEAD = 16, F_ANY_DELAY = 32, F_SUB_ON = 64 };
enum : uint8_t { O_ENABLED = 1, O_MUTE = 2, O_PAIR_OFF = 4 };

struct ChainQ {
    uint32_t N, N_pad, nb, max_frames;
    int32_t *bq;                                   // [role][band 12][8][N_pad]: b0 b1 b2 a1 a2 s1 s2 bypass
    int32_t *preamp;                               // [2][N_pad]
    uint8_t *flags;
    int32_t *loud_c; int32_t *loud_st; uint8_t *loud_byp;    // [2 j][5][N_pad], [2 side][2 j][2][N_pad], [N_pad]
    int32_t *xf;                                   // [7][N_pad]
    float *lev_c; int32_t *lev_i; float *lev_f; uint32_t *lev_idx; int32_t *lev_la;   // [9][Np], [4][Np] env_l env_r gain gain_prev, [Np] smooth_db, [Np], [2][480][Np]
    int32_t *o_gl, *o_gr, *o_gain; uint8_t *o_flags; int32_t *o_dly;                  // [5][N_pad]
    int32_t *dline; uint32_t *widx_in, *widx_out;  // [5][N_pad][2048], [N_pad]
    int32_t *pdm;                                  // [9][N_pad]
    uint16_t *peaks; uint16_t *clip;               // [7][N_pad], [N_pad]
    int32_t *master; int32_t *subq;                // [2][max_frames][N_pad], [max_frames][N_pad]
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

struct QC { int32_t hi; uint32_t lo, hi16; };
__device__ __forceinline__ QC qsplit(int32_t c) { QC r; r.hi = c >> 16; r.lo = (uint32_t)c & 0xFFFFu; r.hi16 = (uint32_t)r.hi << 4; return r; }
__device__ __forceinline__ uint32_t mulq(const QC &c, int32_t xh, uint32_t xl)
{
    const uint32_t mid = (uint32_t)c.hi * xl + c.lo * (uint32_t)xh;
    return c.hi16 * (uint32_t)xh + (uint32_t)((int32_t)mid >> 12);
}

// dsp_process_channel_block(), dsp_process_rp2040.S:225-394, over a lane-private column xs[i * 32]
__device__ __forceinline__ void eq_packet(const ChainQ &d, uint32_t role, uint32_t inst, int32_t *xs, uint32_t n)
{
    const size_t Np = d.N_pad;
    for (uint32_t b = 0; b < d.nb; b++) {
        int32_t *base = d.bq + ((size_t)(role * DSPI_MAX_BANDS + b) * 8) * Np + inst;
        if (base[7 * Np]) continue;                                           // bypass byte, .S:246-248
        const QC c0 = qsplit(base[0]), c1 = qsplit(base[1 * Np]), c2 = qsplit(base[2 * Np]), c3 = qsplit(base[3 * Np]), c4 = qsplit(base[4 * Np]);
        uint32_t s1 = (uint32_t)base[5 * Np], s2 = (uint32_t)base[6 * Np];
#pragma unroll 4
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t x = (uint32_t)xs[i * 32];
            const int32_t xh = (int32_t)x >> 16;
            const uint32_t xl = x & 0xFFFFu;
            const uint32_t y = mulq(c0, xh, xl) + s1;                         // :273-285
            const uint32_t t1 = mulq(c1, xh, xl), t3 = mulq(c2, xh, xl);      // :288-312
            const int32_t yh = (int32_t)y >> 16;
            const uint32_t yl = y & 0xFFFFu;
            const uint32_t t2 = mulq(c3, yh, yl), t4 = mulq(c4, yh, yl);      // :319-348
            s1 = (t1 - t2) + s2;                                              // :332-335
            s2 = t3 - t4;                                                     // :351-353
            xs[i * 32] = (int32_t)y;
        }
        base[5 * Np] = (int32_t)s1;
        base[6 * Np] = (int32_t)s2;
    }
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
__global__ void __launch_bounds__(128, 2)
chainq_front_kernel(ChainQ d, const uint8_t *__restrict__ pcm, uint32_t bit_depth, uint32_t p0, uint32_t n_packets, uint32_t fpp, uint32_t F)
{
    extern __shared__ int32_t smem_q[];                    // [warps][kPkt][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t side = lane >> 4;
    const uint32_t wbase = (blockIdx.x * (blockDim.x >> 5) + warp) * 16;
    if (wbase >= d.N_pad) return;
    const uint32_t inst = wbase + (lane & 15);
    const bool live = inst < d.N;
    int32_t *xs = smem_q + (size_t)warp * kPkt * 32 + lane;
    const size_t Np = d.N_pad;

    const uint8_t flags = d.flags[inst];
    const bool loud_on = flags & F_LOUD, lev_on = flags & F_LEV, xf_on = flags & F_XFEED;
    const bool skip_master = flags & F_BYPASS_MASTER, lookahead = flags & F_LOOKAHEAD;
    const int32_t preamp = d.preamp[side * Np + inst];

    int32_t lc[2][5], ls[2][2];
    const uint8_t loud_byp = d.loud_byp[inst];
#pragma unroll
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int k = 0; k < 5; k++) lc[j][k] = d.loud_c[(j * 5 + k) * Np + inst];
        ls[j][0] = d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst];
        ls[j][1] = d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst];
    }
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

    const uint32_t bpf = bit_depth == 24 ? 6u : 4u;
    const uint8_t *my_pcm = pcm + ((size_t)inst * F) * bpf + side * (bpf / 2);
    int32_t peak_last = 0;
    uint16_t clip = 0;

    for (uint32_t p = p0; p < p0 + n_packets; p++) {
        const uint32_t f0 = p * fpp;
        // PASS 1 (:997-1015) + loudness (:1018-1047)
        for (uint32_t i = 0; i < fpp; i++) {
            int32_t raw = 0;
            if (live) {
                const uint8_t *q = my_pcm + (size_t)(f0 + i) * bpf;
                if (bit_depth == 24) raw = ((int32_t)((uint32_t)q[2] << 24 | (uint32_t)q[1] << 16 | (uint32_t)q[0] << 8)) >> 2;    // :1001
                else raw = (int32_t)((uint32_t)(int32_t)(int16_t)((uint16_t)q[0] | (uint16_t)q[1] << 8) << 14);                   // :1010
            }
            int32_t x = mul_q28(raw, preamp);
            if (loud_on) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if ((loud_byp >> j) & 1) continue;
                    const int32_t result = mul_q28(lc[j][0], x) + ls[j][0];                              // :1026
                    ls[j][0] = mul_q28(lc[j][1], x) - mul_q28(lc[j][3], result) + ls[j][1];
                    ls[j][1] = mul_q28(lc[j][2], x) - mul_q28(lc[j][4], result);
                    x = result;
                }
            }
            xs[i * 32] = x;
        }
        // PASS 2 (:1050-1055)
        if (!skip_master) eq_packet(d, side, inst, xs, fpp);
        __syncwarp();

        // PASS 2.5: leveller (leveller.c:275-389); every lane walks the same shuffles
        if (__any_sync(0xffffffffu, lev_on)) {
            const int32_t a_rms = __float2int_rz(__fmul_rn(lvc[0], 268435456.0f));                       // :286
            const int32_t one_minus = kUnity - a_rms;
            int32_t e = env;
            for (uint32_t i = 0; i < fpp; i++) {                                                         // :292-299
                const int32_t s = xs[i * 32];
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
            for (uint32_t i = 0; i < fpp; i++) {                                                          // :347-386
                int32_t gain;
                if (fpp == 1) gain = g_cur;
                else gain = g_prev + (int32_t)(((int64_t)(g_cur - g_prev) * (int64_t)i) / (int32_t)(fpp - 1));   // :352
                int32_t o = xs[i * 32];
                if (lev_on && lookahead) {
                    const int32_t held = la_buf[(size_t)la_idx * Np];
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
                if (lev_on) xs[i * 32] = mul_q28(o, gain);
            }
            if (lev_on) {
                env = e;
                smooth_db = new_smooth;
                gain_prev_q = g_prev;
                gain_q = g_cur;
            }
        }

        // PASS 3 (:1065-1073)
        int32_t pk = 0;
        int32_t *mout = d.master + ((size_t)side * d.max_frames + f0) * Np + inst;
        for (uint32_t i = 0; i < fpp; i++) {
            int32_t v = xs[i * 32];
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
            mout[(size_t)i * Np] = v;
        }
        peak_last = pk;
        if (pk > kClipThresh) clip |= (uint16_t)(1u << side);
        __syncwarp();
    }

#pragma unroll
    for (int j = 0; j < 2; j++) {
        d.loud_st[((side * 2 + j) * 2 + 0) * Np + inst] = ls[j][0];
        d.loud_st[((side * 2 + j) * 2 + 1) * Np + inst] = ls[j][1];
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
__global__ void __launch_bounds__(128, 2)
chainq_out_kernel(ChainQ d, uint32_t p0, uint32_t n_packets, uint32_t fpp, uint32_t F, int32_t *__restrict__ spdif_out)
{
    extern __shared__ int32_t smem_q[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + warp;
    const uint32_t groups = d.N_pad / 32;
    if (wid >= groups * kOuts) return;
    const uint32_t o = wid / groups;
    const uint32_t inst = (wid % groups) * 32 + lane;
    const bool live = inst < d.N;
    const size_t Np = d.N_pad;
    int32_t *ys = smem_q + (size_t)warp * kPkt * 32 + lane;

    const uint8_t of = d.o_flags[o * Np + inst];
    const bool enabled = of & O_ENABLED, mute = of & O_MUTE, pair_off = of & O_PAIR_OFF;
    const int32_t gl = d.o_gl[o * Np + inst], gr = d.o_gr[o * Np + inst], gain = d.o_gain[o * Np + inst];
    const int32_t dly = d.o_dly[o * Np + inst];
    const uint8_t iflags = d.flags[inst];
    const bool any_delay = iflags & F_ANY_DELAY, delay_on = any_delay && dly > 0;
    const bool run_eq = enabled && !mute && !(iflags & F_BYPASS_MASTER);                                  // :1197-1201 (quirk: gated on bypass_master_eq)
    uint32_t widx = d.widx_in[inst];
    int32_t *ring = d.dline + ((size_t)o * Np + inst) * kMaxDelay;
    const bool is_sub = o == kOuts - 1;
    int32_t *my_spdif = nullptr;
    if (!is_sub && spdif_out && live) my_spdif = spdif_out + (((size_t)inst * 2 + (o >> 1)) * F) * 2 + (o & 1);

    int32_t peak_last = 0;
    uint16_t clip = 0;
    for (uint32_t p = p0; p < p0 + n_packets; p++) {
        const uint32_t f0 = p * fpp;
        // PASS 4 (:1076-1100)
        for (uint32_t i = 0; i < fpp; i++) {
            int32_t v = 0;
            if (enabled) {
                const int32_t l = d.master[((size_t)0 * d.max_frames + f0 + i) * Np + inst];
                const int32_t r = d.master[((size_t)1 * d.max_frames + f0 + i) * Np + inst];
                if (gl != 0 && gr != 0) v = mul_q15(l, gl) + mul_q15(r, gr);
                else if (gl != 0) v = mul_q15(l, gl);
                else if (gr != 0) v = mul_q15(r, gr);
            }
            ys[i * 32] = v;
        }
        // PASS 5 (:1196-1213)
        if (run_eq) eq_packet(d, 2 + o, inst, ys, fpp);
        int32_t pk = 0;
        for (uint32_t t0 = 0; t0 < fpp; t0 += 8) {
            const int nvalid = min(8, (int)(fpp - t0));
            int32_t x[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[i] = (i < nvalid) ? ys[(t0 + i) * 32] : 0;
                if (enabled) x[i] = (gain == 0) ? 0 : mul_q15(x[i], gain);                                // :1206-1212
            }
            if (delay_on) {                                                                               // PASS 6 (:1216-1230)
                if (dly <= kMaxDelay - 8) {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if (i < nvalid) ring[(widx + t0 + i) & (kMaxDelay - 1)] = x[i];
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if (i < nvalid) x[i] = ring[(widx + t0 + i - (uint32_t)dly) & (kMaxDelay - 1)];
                } else {
                    for (int i = 0; i < nvalid; i++) {
                        ring[(widx + t0 + i) & (kMaxDelay - 1)] = x[i];
                        x[i] = ring[(widx + t0 + i - (uint32_t)dly) & (kMaxDelay - 1)];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i >= nvalid) break;
                const int32_t v = x[i];
                const int32_t a = abs(v);
                if (a > pk) pk = a;
                if (is_sub) {
                    if (enabled) d.subq[(size_t)(f0 + t0 + i) * Np + inst] = v;                           // :1270 pdm_push_sample(buf_out[pdm_out][i])
                } else if (my_spdif) {
                    int32_t word = 0;
                    if (!pair_off) {
                        word = (v + 32) >> 6;                                                            // :1254-1255
                        word = word > 0x7FFFFF ? 0x7FFFFF : (word < -0x800000 ? -0x800000 : word);        // clip_s24, config.h:547-551
                    }
                    my_spdif[(size_t)(f0 + t0 + i) * 2] = word;
                }
            }
        }
        if (any_delay) widx = (widx + fpp) & (kMaxDelay - 1);
        peak_last = pk;
        if (pk > kClipThresh && (!is_sub || enabled)) clip |= 1;
        __syncwarp();
    }
    uint16_t pq = (uint16_t)(peak_last >> 13);                                                            // :1239 / :1267
    if (is_sub && !enabled) pq = 0;                                                                       // :1273
    d.peaks[(2 + o) * Np + inst] = pq;
    if (clip) atomicOr(reinterpret_cast<unsigned int *>(d.clip + (inst & ~1u)), (1u << (2 + o)) << (16 * (inst & 1)));
    if (o == 0) d.widx_out[inst] = widx;
}

__global__ void __launch_bounds__(64)
chainq_pdm_kernel(ChainQ d, uint32_t f_begin, uint32_t f_end, uint32_t F, uint32_t *__restrict__ pdm_out)
{
    const uint32_t inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= d.N) return;
    if (!(d.flags[inst] & F_SUB_ON)) return;                                                              // usb_audio.c:1261
    pdm_modulate_frames(d.pdm, d.subq + inst, d.N_pad, d.N_pad, inst, f_begin, f_end, F, pdm_out);
}

// filters[][] of n instances (instance-major AoS, 32-byte records) <-> [role][band][8][N_pad]
__global__ void chainq_pack_kernel(const dspi_biquad_q28 *__restrict__ aos, uint32_t inst0, uint32_t n, ChainQ d)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kRoles * DSPI_MAX_BANDS) return;
    const uint32_t inst = inst0 + i / (kRoles * DSPI_MAX_BANDS), rb = i % (kRoles * DSPI_MAX_BANDS);
    const dspi_biquad_q28 &q = aos[(size_t)inst * kRoles * DSPI_MAX_BANDS + rb];
    int32_t *dst = d.bq + ((size_t)rb * 8) * d.N_pad + inst;
    const size_t Np = d.N_pad;
    dst[0] = q.b0; dst[1 * Np] = q.b1; dst[2 * Np] = q.b2; dst[3 * Np] = q.a1; dst[4 * Np] = q.a2;
    dst[5 * Np] = q.s1; dst[6 * Np] = q.s2; dst[7 * Np] = q.bypass ? 1 : 0;
}
__global__ void chainq_unpack_kernel(dspi_biquad_q28 *__restrict__ aos, uint32_t inst0, uint32_t n, ChainQ d)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kRoles * DSPI_MAX_BANDS) return;
    const uint32_t inst = inst0 + i / (kRoles * DSPI_MAX_BANDS), rb = i % (kRoles * DSPI_MAX_BANDS);
    dspi_biquad_q28 &q = aos[(size_t)inst * kRoles * DSPI_MAX_BANDS + rb];
    const int32_t *src = d.bq + ((size_t)rb * 8) * d.N_pad + inst;
    q.s1 = src[5 * (size_t)d.N_pad];
    q.s2 = src[6 * (size_t)d.N_pad];
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
    std::vector<void *> allocs;
    uint64_t launches;
    void *d_pcm; size_t pcm_bytes;
    int32_t *d_spdif; size_t spdif_bytes;
    uint32_t *d_pdmout; size_t pdmout_bytes;
    dspi_status_q28 *d_status;
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
    for (void *p : c->allocs) cudaFree(p);
    if (c->d_pcm) cudaFree(c->d_pcm);
    if (c->d_spdif) cudaFree(c->d_spdif);
    if (c->d_pdmout) cudaFree(c->d_pdmout);
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
    c->d_aos = nullptr; c->launches = 0; c->d_pcm = nullptr; c->pcm_bytes = 0; c->d_spdif = nullptr; c->spdif_bytes = 0;
    c->d_pdmout = nullptr; c->pdmout_bytes = 0; c->d_status = nullptr;
    c->desc = *desc;
    ChainQ &d = c->d;
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
    TRY(dev_alloc(c, &d.bq, Np * dspi::kRoles * DSPI_MAX_BANDS * 8));
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
    TRY(dev_alloc(c, &d.master, (size_t)2 * d.max_frames * Np, false));
    TRY(dev_alloc(c, &d.subq, (size_t)d.max_frames * Np, false));
    TRY(dev_alloc(c, &c->d_status, Np));
    // every band of every channel starts bypassed (dsp_init_default_filters, dsp_pipeline.c:177-199)
    if (e == cudaSuccess) {
        std::vector<int32_t> ones(Np, 1);
        for (int rb = 0; rb < dspi::kRoles * DSPI_MAX_BANDS && e == cudaSuccess; rb++)
            e = cudaMemcpyAsync(d.bq + ((size_t)rb * 8 + 7) * Np, ones.data(), Np * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
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
    std::vector<float> lev_c(9 * n);
    std::vector<uint8_t> flags(n), loud_byp(n), oflags(O * n);
    for (uint32_t i = 0; i < n; i++) {
        const dspi_chain_params_q28 &p = params[i];
        int32_t vol_mul = p.host_mute ? 0 : (int32_t)p.host_vol_mul;                                    // usb_audio.c:975
        int32_t pmg = (int32_t)(p.preset_mute_gain * 32768.0f + 0.5f);                                  // :976-978
        if (pmg < 0) pmg = 0;
        if (pmg > 32768) pmg = 32768;
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
            uint8_t f = (oc.enabled ? dspi::O_ENABLED : 0) | (oc.mute ? dspi::O_MUTE : 0);
            if (o < O - 1 && !oc.enabled && !p.matrix.outputs[o ^ 1].enabled) f |= dspi::O_PAIR_OFF;    // :1248-1251
            oflags[o * n + i] = f;
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
        loud_byp[i] = (p.loudness[0].bypass ? 1 : 0) | (p.loudness[1].bypass ? 2 : 0);
        for (int j = 0; j < 2; j++) {
            const int32_t v[5] = { p.loudness[j].b0, p.loudness[j].b1, p.loudness[j].b2, p.loudness[j].a1, p.loudness[j].a2 };
            for (int k = 0; k < 5; k++) loud_c[(j * 5 + k) * n + i] = v[k];
        }
        const int32_t xv[7] = { p.crossfeed.lp_a0, p.crossfeed.lp_b1, p.crossfeed.lp_state_L, p.crossfeed.lp_state_R,
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
    CU_OK(put(d.loud_c, loud_c.data(), 10, 4));
    CU_OK(put(d.loud_byp, loud_byp.data(), 1, 1));
    CU_OK(put(d.xf, xf.data(), 7, 4));
    CU_OK(put(d.lev_c, lev_c.data(), 9, 4));
    CU_OK(put(d.o_gl, gl.data(), O, 4));
    CU_OK(put(d.o_gr, gr.data(), O, 4));
    CU_OK(put(d.o_gain, gain.data(), O, 4));
    CU_OK(put(d.o_flags, oflags.data(), O, 1));
    CU_OK(put(d.o_dly, dly.data(), O, 4));
    CU_OK(cudaStreamSynchronize(c->stream));
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
    dspi::chainq_pack_kernel<<<(n * row + 127) / 128, 128, 0, c->stream>>>(c->d_aos, inst0, n, c->d);
    CU_OK(cudaGetLastError());
    c->launches++;
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

int dspi_chainq_download_biquads(dspi_chainq *c, uint32_t inst0, uint32_t n, dspi_biquad_q28 *biquads)
{
    if (!c || !biquads) return fail(DSPI_EINVAL, "null argument");
    if ((uint64_t)inst0 + n > c->desc.n_instances) return fail(DSPI_ERANGE, "instances [%u, %u) outside engine of %u", inst0, inst0 + n, c->desc.n_instances);
    if (n == 0) return DSPI_OK;
    CU_OK(cudaSetDevice(c->desc.device));
    const size_t row = (size_t)dspi::kRoles * DSPI_MAX_BANDS;
    dspi::chainq_unpack_kernel<<<(n * row + 127) / 128, 128, 0, c->stream>>>(c->d_aos, inst0, n, c->d);
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
    const size_t smem = (size_t)4 * dspi::kPkt * 32 * 4;
    static bool configured = false;
    if (!configured) {
        CU_OK(cudaFuncSetAttribute(dspi::chainq_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_OK(cudaFuncSetAttribute(dspi::chainq_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // Stage pipeline over packet slices on three streams (chain_streams.cuh).
    dspi::ChainStreams &st = c->st;
    const uint32_t n_slices = n_packets < (uint32_t)dspi::ChainStreams::kMaxSlices ? n_packets : (uint32_t)dspi::ChainStreams::kMaxSlices;
    CU_OK(cudaEventRecord(st.ev_begin, c->stream));
    CU_OK(cudaStreamWaitEvent(st.s_front, st.ev_begin, 0));
    for (uint32_t sl = 0; sl < n_slices; sl++) {
        const uint32_t p0 = (uint32_t)((uint64_t)n_packets * sl / n_slices), p1 = (uint32_t)((uint64_t)n_packets * (sl + 1) / n_slices);
        const ChainQ d = c->d;
        const uint32_t fwarps = d.N_pad / 16, owarps = d.N_pad / 32 * dspi::kOuts;
        dspi::chainq_front_kernel<<<(fwarps + 3) / 4, 128, smem, st.s_front>>>(d, (const uint8_t *)d_pcm, bit_depth, p0, p1 - p0, fpp, F);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(st.ev_front[sl], st.s_front));
        CU_OK(cudaStreamWaitEvent(st.s_out, st.ev_front[sl], 0));
        dspi::chainq_out_kernel<<<(owarps + 3) / 4, 128, smem, st.s_out>>>(d, p0, p1 - p0, fpp, F, d_spdif);
        CU_OK(cudaGetLastError());
        std::swap(c->d.widx_in, c->d.widx_out);
        CU_OK(cudaEventRecord(st.ev_out[sl], st.s_out));
        CU_OK(cudaStreamWaitEvent(st.s_pdm, st.ev_out[sl], 0));
        dspi::chainq_pdm_kernel<<<(d.N + 63) / 64, 64, 0, st.s_pdm>>>(d, p0 * fpp, p1 * fpp, F, d_pdm);
        CU_OK(cudaGetLastError());
        c->launches += 3;
    }
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
    if (pdm_out && pd_bytes > c->pdmout_bytes) { if (c->d_pdmout) cudaFree(c->d_pdmout); c->d_pdmout = nullptr; c->pdmout_bytes = 0; CU_OK(cudaMalloc((void **)&c->d_pdmout, pd_bytes)); c->pdmout_bytes = pd_bytes; }
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

int dspi_chainq_sync(dspi_chainq *c)
{
    if (!c) return fail(DSPI_EINVAL, "null argument");
    CU_OK(cudaSetDevice(c->desc.device));
    CU_OK(cudaStreamSynchronize(c->stream));
    return DSPI_OK;
}

uint64_t dspi_chainq_launch_count(dspi_chainq *c) { return c ? c->launches : 0; }

}  // extern "C"
