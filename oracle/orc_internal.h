/* orc_internal.h — helpers shared by the oracle translation units (TEST INFRASTRUCTURE). */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include <stdint.h>
#include <math.h>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

extern int orc_x86_cvt_mode;
extern int orc_libm_f64_mode;

/* FTZ (bit 15) | DAZ (bit 6): the firmware sets FPSCR FZ+DN on both cores
 * (main.c:593-600, pdm_generator.c:693-700). */
static inline unsigned orc_ftz_enter(void)
{
#if defined(__x86_64__) || defined(__i386__)
    unsigned old = _mm_getcsr();
    _mm_setcsr(old | 0x8040u);
    return old;
#else
    return 0;
#endif
}
static inline void orc_ftz_leave(unsigned old)
{
#if defined(__x86_64__) || defined(__i386__)
    _mm_setcsr(old);
#else
    (void)old;
#endif
}

/* leveller.c:178,200,206 call log10f/powf once per block.  Flavour 0: glibc's
 * float routines (bit-identical to oracle/_ref on the same host).  Flavour 1:
 * evaluate in double and round once to float — the definition the CUDA path
 * implements (DESIGN.md "libm policy"). */
static inline float orc_log10f(float x)
{
    return orc_libm_f64_mode ? (float)log10((double)x) : log10f(x);
}
static inline float orc_sinf(float x) { return orc_libm_f64_mode ? (float)sin((double)x) : sinf(x); }
static inline float orc_cosf(float x) { return orc_libm_f64_mode ? (float)cos((double)x) : cosf(x); }
static inline float orc_tanf(float x) { return orc_libm_f64_mode ? (float)tan((double)x) : tanf(x); }
static inline float orc_expf(float x) { return orc_libm_f64_mode ? (float)exp((double)x) : expf(x); }
static inline float orc_logf(float x) { return orc_libm_f64_mode ? (float)log((double)x) : logf(x); }
static inline float orc_powf(float a, float b)
{
    return orc_libm_f64_mode ? (float)pow((double)a, (double)b) : powf(a, b);
}

#endif
