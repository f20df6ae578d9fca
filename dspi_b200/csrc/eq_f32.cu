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
// the SVF state alternates sign every sample (see svf_pair()).
#include "eq_kernels.cuh"

namespace dspi {
namespace {

// ---------------------------------------------------------------------------------------
// value types: float (1 channel / lane) or P2 (2 channels / lane, packed f32x2)
//
// P2 is an opaque 64-bit register pair driven with inline PTX: keeping the pair as ONE .b64
// virtual register forces ptxas to hold every sample/state/coefficient packed for the whole
// kernel.  (With P2 + the __ffma2_rn intrinsics the halves are separate 32-bit values and
// ptxas re-packs them with two MOVs around every packed instruction.)
// ---------------------------------------------------------------------------------------
struct P2 { unsigned long long v; };

__device__ __forceinline__ float v_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float v_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float v_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ P2 v_mul(P2 a, P2 b)
{
    P2 r;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ P2 v_add(P2 a, P2 b)
{
    P2 r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ P2 v_fma(P2 a, P2 b, P2 c)
{
    P2 r;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ P2 p2_pack(float lo, float hi)
{
    P2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void p2_unpack(P2 a, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}

template <typename V> __device__ __forceinline__ V v_bits(unsigned long long b);
template <> __device__ __forceinline__ float v_bits<float>(unsigned long long b) { return __uint_as_float((unsigned)b); }
template <> __device__ __forceinline__ P2 v_bits<P2>(unsigned long long b) { P2 r; r.v = b; return r; }
template <typename V> __device__ __forceinline__ V v_set(float x);
template <> __device__ __forceinline__ float v_set<float>(float x) { return x; }
template <> __device__ __forceinline__ P2 v_set<P2>(float x) { return p2_pack(x, x); }

// sign flip on the integer pipe (exact; a flushed-denormal operand is flushed by the consumer)
__device__ __forceinline__ float v_neg(float a) { return __int_as_float(__float_as_int(a) ^ 0x80000000); }
__device__ __forceinline__ P2 v_neg(P2 a) { P2 r; r.v = a.v ^ 0x8000000080000000ull; return r; }

// ptxas (12.9) contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even with --fmad=false and
// even when the product is written as fma(a, b, -0.0) with a literal -0.0 (it folds that back to
// a multiply first).  The strict flavour therefore forms packed products as fma(a, b, nz) where
// nz = (-0.0, -0.0) arrives as a KERNEL PARAMETER: an exact product rounding (x + -0 == x for
// every x, including both zeros) that the assembler cannot prove foldable.  Scalar FMUL/FADD
// and every fused-flavour sequence are left alone by ptxas (checked in the SASS).
template <bool FUSED> __device__ __forceinline__ float mulx(float a, float b, float) { return __fmul_rn(a, b); }
template <bool FUSED> __device__ __forceinline__ P2 mulx(P2 a, P2 b, P2 nz)
{
    if constexpr (FUSED) return v_mul(a, b);
    else return v_fma(a, b, nz);
}

// a*b + c: one rounding (FUSED) or two (strict)
template <bool FUSED, typename V>
__device__ __forceinline__ V madd(V a, V b, V c, V nz)
{
    if constexpr (FUSED) return v_fma(a, b, c);
    else return v_add(mulx<false>(a, b, nz), c);
}
template <bool FUSED>
__device__ __forceinline__ float madd(float a, float b, float c)
{
    if constexpr (FUSED) return __fmaf_rn(a, b, c);
    else return __fadd_rn(__fmul_rn(a, b), c);
}

// ---------------------------------------------------------------------------------------
// per-band inner loops over a register tile x[N]
// ---------------------------------------------------------------------------------------

// TDF2 biquad, dsp_pipeline.c:354-360.  c = {b0, b1, b2, -a1, -a2}
//   out = b0*in + s1;  s1 = b1*in - a1*out + s2;  s2 = b2*in - a2*out
template <bool FUSED, int N, typename V>
__device__ __forceinline__ void tdf2_tile(const V (&x)[N], V (&y)[N], const V (&c)[6], V &s1, V &s2, const V nz)
{
#pragma unroll
    for (int i = 0; i < N; i++) {
        const V in = x[i];
        const V out = madd<FUSED>(c[0], in, s1, nz);
        const V m = mulx<FUSED>(c[3], out, nz);          // -(a1*out), exact negation of the reference's product
        s1 = v_add(madd<FUSED>(c[1], in, m, nz), s2);
        const V n = mulx<FUSED>(c[4], out, nz);
        s2 = madd<FUSED>(c[2], in, n, nz);
        y[i] = out;
    }
}

// Cytomic SVF, dsp_pipeline.c:299-342.  c = {a1, a2, a3, m0, m1, m2}
//   v3 = in - ic2;  v1 = a1*ic1 + a2*v3;  v2 = ic2 + a2*ic1 + a3*v3;
//   ic1 = 2*v1 - ic1;  ic2 = 2*v2 - ic2;  out = mix(in, v1, v2)
// The update `ic = 2v - ic` flips the sign the state enters with, so two samples are
// processed per step: the first with (ic1, ic2), leaving (-ic1', -ic2'); the second consumes
// the negated state and leaves it positive again.  Every rewritten operation is the
// reference's operation with operands negated in pairs, which commutes with rounding.
enum { kMixLP = 2, kMixHP = 3, kMixPK = 4, kMixSH = 5 };

template <bool FUSED, int MIX, int N, typename V>
__device__ __forceinline__ void svf_tile(const V (&x)[N], V (&y)[N], const V (&c)[6], V &ic1, V &ic2, const V nz)
{
    static_assert(N % 2 == 0, "SVF tile processes sample pairs");
    const V kN1 = v_set<V>(-1.0f), kN2 = v_set<V>(-2.0f), kP2 = v_set<V>(2.0f);
    const V na1 = v_neg(c[0]), na3 = v_neg(c[2]), nm2 = v_neg(c[5]);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        {   // ---- state positive on entry, negated on exit
            const V in = x[i];
            const V v3 = v_fma(ic2, kN1, in);                       // in - ic2 (exact product)
            const V p = mulx<FUSED>(c[1], v3, nz);
            V t, v1, v2;
            if constexpr (FUSED) {
                t = v_fma(c[1], ic1, ic2);                          // a2*ic1 + ic2
                v1 = v_fma(c[0], ic1, p);                           // a1*ic1 + a2*v3
                v2 = v_fma(c[2], v3, t);
            } else {
                t = v_add(ic2, mulx<FUSED>(c[1], ic1, nz));
                v1 = v_add(mulx<FUSED>(c[0], ic1, nz), p);
                v2 = v_add(t, mulx<FUSED>(c[2], v3, nz));
            }
            ic1 = v_fma(v1, kN2, ic1);                              // -(2*v1 - ic1)
            ic2 = v_fma(v2, kN2, ic2);
            if constexpr (MIX == kMixLP) y[i] = v2;
            else if constexpr (MIX == kMixPK) y[i] = madd<FUSED>(c[4], v1, in, nz);
            else if constexpr (MIX == kMixHP) y[i] = v_fma(v2, kN1, madd<FUSED>(c[4], v1, in, nz));
            else {
                const V q = mulx<FUSED>(c[4], v1, nz);
                y[i] = madd<FUSED>(c[5], v2, madd<FUSED>(c[3], in, q, nz), nz);
            }
        }
        {   // ---- state negated on entry (n1 = -ic1, n2 = -ic2), positive on exit
            const V in = x[i + 1];
            const V v3 = v_add(in, ic2);                            // in - ic2
            const V p = mulx<FUSED>(c[1], v3, nz);
            V nt, v1, nv2;
            if constexpr (FUSED) {
                nt = v_fma(c[1], ic1, ic2);                         // -(a2*ic1 + ic2)
                v1 = v_fma(na1, ic1, p);                            // a1*ic1 + a2*v3
                nv2 = v_fma(na3, v3, nt);                           // -v2
            } else {
                nt = v_add(ic2, mulx<FUSED>(c[1], ic1, nz));
                v1 = v_add(mulx<FUSED>(na1, ic1, nz), p);
                nv2 = v_add(nt, mulx<FUSED>(na3, v3, nz));
            }
            ic1 = v_fma(v1, kP2, ic1);                              // 2*v1 - ic1
            ic2 = v_fma(nv2, kN2, ic2);                             // 2*v2 - ic2
            if constexpr (MIX == kMixLP) y[i + 1] = mulx<FUSED>(nv2, kN1, nz);
            else if constexpr (MIX == kMixPK) y[i + 1] = madd<FUSED>(c[4], v1, in, nz);
            else if constexpr (MIX == kMixHP) y[i + 1] = v_add(madd<FUSED>(c[4], v1, in, nz), nv2);
            else {
                const V q = mulx<FUSED>(c[4], v1, nz);
                y[i + 1] = madd<FUSED>(nm2, nv2, madd<FUSED>(c[3], in, q, nz), nz);
            }
        }
    }
}

// Scalar, runtime-length, per-lane-mode version: used for warps whose channels do not share
// a band's topology and for the tail of a launch (T not a multiple of the register tile).
// Same operation sequences as above in their natural (reference) form.
template <bool FUSED>
__device__ __noinline__ float2 slow_band(float *xs, int n, uint32_t mode, float c0, float c1, float c2, float c3, float c4, float c5,
                                         float st0, float st1)
{
    if (mode == kModeTdf2) {
        for (int i = 0; i < n; i++) {
            const float in = xs[i];
            const float out = madd<FUSED>(c0, in, st0);
            const float m = __fmul_rn(c3, out);
            st0 = __fadd_rn(madd<FUSED>(c1, in, m), st1);
            const float nn = __fmul_rn(c4, out);
            st1 = madd<FUSED>(c2, in, nn);
            xs[i] = out;
        }
    } else if (mode >= kModeSvfLP) {
        for (int i = 0; i < n; i++) {
            const float in = xs[i];
            const float v3 = __fadd_rn(in, -st1);
            const float p = __fmul_rn(c1, v3);
            float t, v1, v2;
            if (FUSED) {
                t = __fmaf_rn(c1, st0, st1);
                v1 = __fmaf_rn(c0, st0, p);
                v2 = __fmaf_rn(c2, v3, t);
            } else {
                t = __fadd_rn(st1, __fmul_rn(c1, st0));
                v1 = __fadd_rn(__fmul_rn(c0, st0), p);
                v2 = __fadd_rn(t, __fmul_rn(c2, v3));
            }
            st0 = __fmaf_rn(2.0f, v1, -st0);
            st1 = __fmaf_rn(2.0f, v2, -st1);
            float y;
            if (mode == kModeSvfLP) y = v2;
            else if (mode == kModeSvfPK) y = madd<FUSED>(c4, v1, in);
            else if (mode == kModeSvfHP) y = __fadd_rn(madd<FUSED>(c4, v1, in), -v2);
            else y = madd<FUSED>(c5, v2, madd<FUSED>(c3, in, __fmul_rn(c4, v1)));
            xs[i] = y;
        }
    }
    return make_float2(st0, st1);
}

// ---------------------------------------------------------------------------------------
// lane <-> register-tile plumbing
// ---------------------------------------------------------------------------------------
template <typename V> struct Lanes;
template <> struct Lanes<float> {
    static constexpr int CPL = 1;
    __device__ static __forceinline__ float get(float v, int) { return v; }
};
// build a value from per-half scalars (h = 0: channel `lane`, h = 1: channel `lane + 32`)
__device__ __forceinline__ void v_make(float &v, const float (&part)[1]) { v = part[0]; }
__device__ __forceinline__ void v_make(P2 &v, const float (&part)[2]) { v = p2_pack(part[0], part[1]); }
template <> struct Lanes<P2> {
    static constexpr int CPL = 2;
    __device__ static __forceinline__ float get(P2 v, int h) { float lo, hi; p2_unpack(v, lo, hi); return h ? hi : lo; }
};

constexpr int kTileT = 32;          // samples per shared-memory tile row: 128 B == swizzle span
constexpr int kSub = 8;             // samples per register tile
constexpr int kStages = 3;

template <typename V, bool FUSED, int NB>
__global__ void __launch_bounds__(256 * 2 / Lanes<V>::CPL, 1)
eq_f32_kernel(const __grid_constant__ CUtensorMap tmap, float *__restrict__ samples, uint32_t ld, V *__restrict__ coef,
              const uint64_t *__restrict__ modes, uint32_t n_groups, uint32_t n_rows, uint32_t T, uint32_t nb_active, uint32_t use_tma, uint32_t dbg, unsigned long long nz_bits)
{
    constexpr int CPL = Lanes<V>::CPL;
    constexpr int kRows = 32 * CPL;
    constexpr int kWarps = 16 / CPL;
    constexpr uint32_t kStageBytes = kRows * kTileT * 4;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[kWarps][kStages];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x * kWarps + warp;
    if (g >= n_groups) return;                                  // warps are fully independent

    uint8_t *my_smem = smem_raw + (size_t)warp * kStages * kStageBytes;
    uint64_t *full = bars[warp];
    if (lane == 0) {
        if (use_tma) prefetch_tmap(&tmap);
        for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    const V nz = v_bits<V>(nz_bits);                            // (-0.0, -0.0): see mulx()
    const int c0 = g * kRows;                                   // first channel (row) of this group
    const uint32_t ntiles = (T + kTileT - 1) / kTileT;

    auto issue_load = [&](uint32_t tile) {                      // lane 0 only
        const uint32_t s = tile % kStages;
        mbar_arrive_expect_tx(&full[s], kStageBytes);
        tma_load_2d(my_smem + s * kStageBytes, &tmap, &full[s], tile * kTileT, c0);
    };
    const bool mem_on = !(dbg & 2u);                            // diagnostics: DSPI_DBG=2 runs the arithmetic without HBM traffic
    if (use_tma && mem_on && lane == 0)
        for (uint32_t s = 0; s + 1 < kStages && s < ntiles; s++) issue_load(s);

    // ---- coefficients, state and modes of every band -> registers ---------------------
    V c[NB][6], st[NB][2];
    const V *cg = coef + (size_t)g * kMaxBands * 8 * 32;
#pragma unroll
    for (int b = 0; b < NB; b++) {
#pragma unroll
        for (int k = 0; k < 6; k++) c[b][k] = cg[(b * 8 + k) * 32 + lane];
        st[b][0] = cg[(b * 8 + 6) * 32 + lane];
        st[b][1] = cg[(b * 8 + 7) * 32 + lane];
    }
    uint64_t mode_h[CPL];
#pragma unroll
    for (int h = 0; h < CPL; h++) {
        mode_h[h] = modes[(size_t)g * kRows + h * 32 + lane];
        if (nb_active < 16) mode_h[h] &= (1ull << (4 * nb_active)) - 1;          // bands >= nb_active are not processed
    }
    const uint64_t mode_w = __shfl_sync(0xffffffffu, mode_h[0], 0);     // warp-uniform candidate
    uint32_t uni = 0;                                                   // bit b: band b has one topology in this warp
#pragma unroll
    for (int b = 0; b < NB; b++) {
        bool same = true;
#pragma unroll
        for (int h = 0; h < CPL; h++) same = same && (((mode_h[h] ^ mode_w) >> (4 * b)) & 15) == 0;
        if (__all_sync(0xffffffffu, same)) uni |= 1u << b;
    }

    // all NB bands active and TDF2 in every lane -> straight-line path
    uint64_t want = 0;
#pragma unroll
    for (int b = 0; b < NB; b++) want |= (uint64_t)kModeTdf2 << (4 * b);
    bool mine = nb_active == NB;
#pragma unroll
    for (int h = 0; h < CPL; h++) mine = mine && ((mode_h[h] & ((1ull << (4 * NB)) - 1)) == want);
    const bool all_tdf2 = __all_sync(0xffffffffu, mine);

    // ---- stream the tiles ----------------------------------------------------------------
    const uint32_t sw = (lane & 7) << 4;                        // 128B-swizzle XOR for this lane's rows
    for (uint32_t tile = 0; tile < ntiles; tile++) {
        const uint32_t s = tile % kStages;
        uint8_t *buf = my_smem + s * kStageBytes;
        if (use_tma) {
            if (mem_on) mbar_wait(&full[s], (tile / kStages) & 1);
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
#pragma unroll 1
        for (int sub = 0; sub < kTileT / kSub; sub++) {
            const int nvalid = min(kSub, tile_valid - sub * kSub);
            if (nvalid <= 0) break;
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

            if (dbg & 1u) {
                // DSPI_DBG=1: data path only
            } else if (all_tdf2 && nvalid == kSub) {
                // every band of every channel of this warp is a TDF2 biquad: one straight-line block,
                // no dispatch, the scheduler overlaps the bands (wavefront over band x sample)
                V y[kSub];
#pragma unroll
                for (int b = 0; b < NB; b += 2) {
                    tdf2_tile<FUSED>(x, y, c[b], st[b][0], st[b][1], nz);
                    tdf2_tile<FUSED>(y, x, c[b + 1], st[b + 1][0], st[b + 1][1], nz);
                }
            } else {
                V y[kSub];
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    // ping-pong between x and y so no case has to move its results back
                    V(&in)[kSub] = (b & 1) ? y : x;
                    V(&out)[kSub] = (b & 1) ? x : y;
                    const uint32_t m = (b < (int)nb_active) ? ((uint32_t)(mode_w >> (4 * b)) & 15u) : kModeBypass;
                    const bool fast = (((uni >> b) & 1u) || b >= (int)nb_active) && nvalid == kSub;
                    if (fast) {
                        if (m == kModeTdf2) tdf2_tile<FUSED>(in, out, c[b], st[b][0], st[b][1], nz);
                        else if (m == kModeSvfPK) svf_tile<FUSED, kMixPK>(in, out, c[b], st[b][0], st[b][1], nz);
                        else if (m == kModeSvfSH) svf_tile<FUSED, kMixSH>(in, out, c[b], st[b][0], st[b][1], nz);
                        else if (m == kModeSvfLP) svf_tile<FUSED, kMixLP>(in, out, c[b], st[b][0], st[b][1], nz);
                        else if (m == kModeSvfHP) svf_tile<FUSED, kMixHP>(in, out, c[b], st[b][0], st[b][1], nz);
                        else {                                      // bypassed band: dsp_pipeline.c:288
#pragma unroll
                            for (int i = 0; i < kSub; i++) out[i] = in[i];
                        }
                    } else {
                        float xs[CPL][kSub], ns0[CPL], ns1[CPL];
#pragma unroll
                        for (int h = 0; h < CPL; h++) {
#pragma unroll
                            for (int i = 0; i < kSub; i++) xs[h][i] = Lanes<V>::get(in[i], h);
                            const uint32_t mh = (uint32_t)(mode_h[h] >> (4 * b)) & 15u;
                            const float2 ns = slow_band<FUSED>(xs[h], nvalid, mh, Lanes<V>::get(c[b][0], h), Lanes<V>::get(c[b][1], h),
                                                               Lanes<V>::get(c[b][2], h), Lanes<V>::get(c[b][3], h), Lanes<V>::get(c[b][4], h),
                                                               Lanes<V>::get(c[b][5], h), Lanes<V>::get(st[b][0], h), Lanes<V>::get(st[b][1], h));
                            ns0[h] = ns.x;
                            ns1[h] = ns.y;
                        }
                        v_make(st[b][0], ns0);
                        v_make(st[b][1], ns1);
#pragma unroll
                        for (int i = 0; i < kSub; i++) {
                            float part[CPL];
#pragma unroll
                            for (int h = 0; h < CPL; h++) part[h] = xs[h][i];
                            v_make(out[i], part);
                        }
                    }
                }
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

        if (use_tma) {
            fence_proxy_async_smem();                           // my smem writes -> async proxy
            __syncwarp();
            if (lane == 0 && mem_on) {
                tma_store_2d(&tmap, buf, tile * kTileT, c0);
                tma_store_commit();
                const uint32_t nxt = tile + kStages - 1;        // refill the buffer stored one iteration ago
                if (nxt < ntiles) {
                    tma_store_wait_read<1>();
                    issue_load(nxt);
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

    // ---- filter state back to the coefficient store ---------------------------------------
    V *cgw = coef + (size_t)g * kMaxBands * 8 * 32;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        cgw[(b * 8 + 6) * 32 + lane] = st[b][0];
        cgw[(b * 8 + 7) * 32 + lane] = st[b][1];
    }
    if (use_tma && lane == 0) tma_store_wait_all<0>();          // smem must outlive the bulk reads
}

template <typename V, bool FUSED, int NB>
cudaError_t launch_one(const EqLaunch &a, cudaStream_t stream)
{
    constexpr int CPL = Lanes<V>::CPL;
    constexpr int kWarps = 16 / CPL;
    constexpr size_t smem = (size_t)kWarps * kStages * (32 * CPL) * kTileT * 4;
    auto kern = eq_f32_kernel<V, FUSED, NB>;
    static bool configured = false;                             // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const uint32_t n_groups = a.n_groups;
    const uint32_t grid = (n_groups + kWarps - 1) / kWarps;
    kern<<<grid, kWarps * 32, smem, stream>>>(a.tmap, (float *)a.samples, a.ld, (V *)a.coef, a.modes, n_groups, a.n_rows, a.T, a.n_bands, a.use_tma, a.dbg, 0x8000000080000000ull);
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
