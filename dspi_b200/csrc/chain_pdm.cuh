// chain_pdm.cuh — the 2nd-order error-feedback delta-sigma PDM modulator shared by the float and the
// Q28 chain (the firmware feeds it Q28 samples on both platforms: usb_audio.c:953 / :1270).
// Reference: pdm_generator.c:351-397 (steady-state branch: hardware running, a sample available, no
// fade-out), xorshift32 :62-68, noise-shaped TPDF dither :89-108, constants config.h:59-75.
#pragma once
#include <stdint.h>

namespace dspi {

// Instances per thread.  One instance is ONE serial chain (256 dependent decisions per frame): a warp that
// carries one instance per lane issues an instruction every ~2.2 cycles (ncu: 0.45 IPC, stall reason `wait`).
// Two independent instances per lane interleave two such chains in the same instruction stream, which fills
// the idle issue slots: the same warp does twice the work in about the same time.
constexpr int kPdmPerThread = 2;

// state words (SoA, [9][Np]): err1 err2 x1 x2 y1 y2 err_acc rng fade_in_pos.
// Modulates frames [f_begin, f_end) of the K instances inst[k] (active[k] false: that slot is idle - nothing of it
// is read or written): Q28 samples at subq[k][f * frame_stride], 8 words (256 bits, MSB first) per frame to
// pdm_out[(inst[k] * F + f) * 8 ..].
template <int K>
__device__ __forceinline__ void pdm_modulate_frames(int32_t *__restrict__ pdm, const int32_t *const (&subq)[K], size_t frame_stride, uint32_t Np,
                                                    const uint32_t (&inst)[K], const bool (&active)[K], uint32_t f_begin, uint32_t f_end, uint32_t F,
                                                    uint32_t *__restrict__ pdm_out)
{
    int32_t err1[K], err2[K], x1[K], x2[K], y1[K], y2[K], err_acc[K], q_next[K];
    uint32_t rng[K], fade[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        err1[k] = err2[k] = x1[k] = x2[k] = y1[k] = y2[k] = err_acc[k] = q_next[k] = 0;
        rng[k] = 1; fade[k] = 1024;
        if (active[k]) {
            const uint32_t i = inst[k];
            err1[k] = pdm[0 * Np + i]; err2[k] = pdm[1 * Np + i];
            x1[k] = pdm[2 * Np + i]; x2[k] = pdm[3 * Np + i]; y1[k] = pdm[4 * Np + i]; y2[k] = pdm[5 * Np + i];
            err_acc[k] = pdm[6 * Np + i];
            rng[k] = (uint32_t)pdm[7 * Np + i]; fade[k] = (uint32_t)pdm[8 * Np + i];
            if (f_begin < f_end) q_next[k] = subq[k][(size_t)f_begin * frame_stride];
        }
    }
    for (uint32_t f = f_begin; f < f_end; f++) {
        int32_t target[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            int32_t pcm = q_next[k] >> 14;                                   // :352
            if (active[k] && f + 1 < f_end) q_next[k] = subq[k][(size_t)(f + 1) * frame_stride];   // next frame's load overlaps this frame's 256 decisions
            pcm = max(-29500, min(29500, pcm));                              // :353-354
            if (fade[k] < 1024u) { pcm = (pcm * (int32_t)fade[k]) >> 10; fade[k]++; }   // :357-360
            target[k] = pcm + 32768;
        }
        // one 32-decision chunk per iteration, NOT unrolled: 2 x 32 x 6 instructions of straight-line code stay inside
        // the instruction cache (scripts/pdm_ubench.cu: beyond ~32 KB of loop body the fetch rate collapses)
#pragma unroll 1
        for (int chunk = 0; chunk < 8; chunk++) {
            int32_t dither[K], s[K], g[K], tg[K];
            uint32_t inv[K];                                                 // complement of the output word
#pragma unroll
            for (int k = 0; k < K; k++) {
                rng[k] ^= rng[k] << 13; rng[k] ^= rng[k] >> 17; rng[k] ^= rng[k] << 5;          // :63-68
                const int32_t raw = (int32_t)(rng[k] & 0x1FFu) - 255;        // :368
                err_acc[k] = ((err_acc[k] * 248) >> 8) + ((err2[k] >> 8) >> 6);                 // :92
                const int32_t in = raw - err_acc[k];
                dither[k] = (15778 * in - 31556 * x1[k] + 15778 * x2[k] + 31531 * y1[k] - 15580 * y2[k]) >> 14;   // :98-99
                x2[k] = x1[k]; x1[k] = in; y2[k] = y1[k]; y1[k] = dither[k];
                // :372-378 restated on two running sums so that only TWO dependent integer ops separate
                // consecutive decisions:
                //   s = err2 + dither   (the comparator input)       g = err1 + target
                //   bit = s >= 0;  s' = s + g - 2*fb;  g' = g + target - fb        (fb = bit ? 65535 : 0)
                // which is err1 += target - fb; err2 += err1 - fb with the substitutions above (all int32,
                // wrapping like the reference).  With m = s >> 31 (0 when the bit is 1, -1 when it is 0) the
                // corrections become m * -K + (sum - K): one shift on the ALU pipe feeding one IMAD on the
                // FMA pipe per decision (scripts/pdm_ubench.cu: 13.5 cycles per decision for one chain per
                // warp against 18.0 for the mask form, 20.4 with predicated corrections, 23.4 for the
                // reference's own statement order).
                inv[k] = 0;
                s[k] = err2[k] + dither[k];
                g[k] = err1[k] + target[k];
                tg[k] = target[k] - 65535;
            }
#pragma unroll
            for (int b = 0; b < 32; b++) {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int32_t m = s[k] >> 31;
                    const int32_t t2 = s[k] + g[k] - 2 * 65535;
                    const int32_t g2 = g[k] + tg[k];
                    inv[k] = __funnelshift_l((uint32_t)s[k], inv[k], 1);     // shift the sign in, MSB first (:375)
                    s[k] = m * (-2 * 65535) + t2;
                    g[k] = m * -65535 + g2;
                }
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                err2[k] = s[k] - dither[k];
                err1[k] = g[k] - target[k];
                if (pdm_out && active[k]) pdm_out[((size_t)inst[k] * F + f) * 8 + chunk] = ~inv[k];   // :380
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            err1[k] -= err1[k] >> 16;                                        // :396-397
            err2[k] -= err2[k] >> 16;
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!active[k]) continue;
        const uint32_t i = inst[k];
        pdm[0 * Np + i] = err1[k]; pdm[1 * Np + i] = err2[k];
        pdm[2 * Np + i] = x1[k]; pdm[3 * Np + i] = x2[k]; pdm[4 * Np + i] = y1[k]; pdm[5 * Np + i] = y2[k];
        pdm[6 * Np + i] = err_acc[k]; pdm[7 * Np + i] = (int32_t)rng[k]; pdm[8 * Np + i] = (int32_t)fade[k];
    }
}

}  // namespace dspi
