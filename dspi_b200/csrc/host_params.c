/*
 * host_params.c — host side of the parameter API: EqParamPacket -> Biquad.
 *
 * Takes the role of dsp_compute_coefficients() (reference
 * firmware/DSPi/dsp_pipeline.c:61-175) for both coefficient stores.  Runs on
 * the host so that coefficients are produced with the host libm and handed to
 * the GPU as data.  Built with -ffp-contract=off: every
 * operation is rounded on its own, the same arithmetic as the reference
 * compiled without contraction.
 */
#include <math.h>
#include <string.h>
#include "dspi_b200.h"

#define PI_F 3.1415926535f          /* the reference's literal, dsp_pipeline.c:97,145 */

/* dsp_pipeline.c:6-17 */
static int recipe_is_flat(const dspi_eq_param *p)
{
    if (p->type == DSPI_FILTER_FLAT || p->freq <= 0.0f) return 1;
    switch (p->type) {
    case DSPI_FILTER_PEAKING: case DSPI_FILTER_LOWSHELF: case DSPI_FILTER_HIGHSHELF:
        return fabsf(p->gain_db) < 0.01f;
    default:
        return 0;
    }
}

/* dsp_pipeline.c:78-81 — the clamps are written back into the caller's recipe */
static void recipe_clamp(dspi_eq_param *p, float fs)
{
    float q = p->Q, f = p->freq;
    if (q < 0.1f) q = 0.1f;
    if (q > 20.0f) q = 20.0f;
    if (f < 10.0f) f = 10.0f;
    if (f > fs * 0.45f) f = fs * 0.45f;
    p->Q = q;
    p->freq = f;
}

/* un-normalised RBJ cookbook section, dsp_pipeline.c:145-156; n = {b0,b1,b2}, d = {a0,a1,a2} */
static void cookbook(const dspi_eq_param *p, float A, float fs, float n[3], float d[3])
{
    const float omega = 2.0f * PI_F * p->freq / fs;
    const float sn = sinf(omega), cs = cosf(omega);
    const float alpha = sn / (2.0f * p->Q);
    n[0] = 1.0f; n[1] = 0.0f; n[2] = 0.0f;
    d[0] = 1.0f; d[1] = 0.0f; d[2] = 0.0f;
    switch (p->type) {
    case DSPI_FILTER_LOWPASS:
        n[0] = (1 - cs) / 2; n[1] = 1 - cs; n[2] = (1 - cs) / 2;
        d[0] = 1 + alpha; d[1] = -2 * cs; d[2] = 1 - alpha;
        break;
    case DSPI_FILTER_HIGHPASS:
        n[0] = (1 + cs) / 2; n[1] = -(1 + cs); n[2] = (1 + cs) / 2;
        d[0] = 1 + alpha; d[1] = -2 * cs; d[2] = 1 - alpha;
        break;
    case DSPI_FILTER_PEAKING:
        n[0] = 1 + alpha * A; n[1] = -2 * cs; n[2] = 1 - alpha * A;
        d[0] = 1 + alpha / A; d[1] = -2 * cs; d[2] = 1 - alpha / A;
        break;
    case DSPI_FILTER_LOWSHELF:
        n[0] = A * ((A + 1) - (A - 1) * cs + 2 * sqrtf(A) * alpha);
        n[1] = 2 * A * ((A - 1) - (A + 1) * cs);
        n[2] = A * ((A + 1) - (A - 1) * cs - 2 * sqrtf(A) * alpha);
        d[0] = (A + 1) + (A - 1) * cs + 2 * sqrtf(A) * alpha;
        d[1] = -2 * ((A - 1) + (A + 1) * cs);
        d[2] = (A + 1) + (A - 1) * cs - 2 * sqrtf(A) * alpha;
        break;
    case DSPI_FILTER_HIGHSHELF:
        n[0] = A * ((A + 1) + (A - 1) * cs + 2 * sqrtf(A) * alpha);
        n[1] = -2 * A * ((A - 1) + (A + 1) * cs);
        n[2] = A * ((A + 1) + (A - 1) * cs - 2 * sqrtf(A) * alpha);
        d[0] = (A + 1) - (A - 1) * cs + 2 * sqrtf(A) * alpha;
        d[1] = 2 * ((A - 1) - (A + 1) * cs);
        d[2] = (A + 1) - (A - 1) * cs - 2 * sqrtf(A) * alpha;
        break;
    default:
        break;
    }
}

void dspi_compute_coefficients_f32(dspi_eq_param *p, dspi_biquad_f32 *bq, float fs)
{
    if (recipe_is_flat(p) || fs == 0) {                 /* dsp_pipeline.c:62-73 */
        bq->bypass = 1;
        bq->use_svf = 0;
        bq->b0 = 1.0f;
        bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0.0f;
        bq->sva1 = bq->sva2 = bq->sva3 = 0.0f;
        bq->svm0 = bq->svm1 = bq->svm2 = 0.0f;
        return;
    }
    bq->bypass = 0;
    recipe_clamp(p, fs);
    const float A = powf(10.0f, p->gain_db / 40.0f);    /* :83 */

    /* :87-92 — pick the topology; a flip clears both state pairs */
    const uint8_t svf_now = (p->freq < (fs / 7.5f)) ? 1 : 0;
    if (svf_now != bq->use_svf) {
        bq->s1 = bq->s2 = 0.0f;
        bq->svic1eq = bq->svic2eq = 0.0f;
    }
    bq->use_svf = svf_now;

    if (svf_now) {                                      /* :94-138 Cytomic SVF */
        float g = tanf(PI_F * p->freq / fs);
        float k = 1.0f / p->Q;
        if (p->type == DSPI_FILTER_PEAKING)        k = 1.0f / (p->Q * A);
        else if (p->type == DSPI_FILTER_LOWSHELF)  g = g / sqrtf(A);
        else if (p->type == DSPI_FILTER_HIGHSHELF) g = g * sqrtf(A);
        const float c1 = 1.0f / (1.0f + g * (g + k));
        const float c2 = g * c1;
        const float c3 = g * c2;
        float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f;
        switch (p->type) {
        case DSPI_FILTER_LOWPASS:   m2 = 1.0f; break;
        case DSPI_FILTER_HIGHPASS:  m0 = 1.0f; m1 = -k; m2 = -1.0f; break;
        case DSPI_FILTER_PEAKING:   m0 = 1.0f; m1 = k * (A * A - 1.0f); break;
        case DSPI_FILTER_LOWSHELF:  m0 = 1.0f; m1 = k * (A - 1.0f); m2 = A * A - 1.0f; break;
        case DSPI_FILTER_HIGHSHELF: m0 = A * A; m1 = k * (1.0f - A) * A; m2 = 1.0f - A * A; break;
        default: break;
        }
        bq->sva1 = c1; bq->sva2 = c2; bq->sva3 = c3;
        bq->svm0 = m0; bq->svm1 = m1; bq->svm2 = m2;
        bq->svf_type = p->type;
        bq->b0 = 1.0f;
        bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0.0f;
        return;
    }

    bq->sva1 = bq->sva2 = bq->sva3 = 0.0f;              /* :141-142 */
    bq->svm0 = bq->svm1 = bq->svm2 = 0.0f;
    float n[3], d[3];
    cookbook(p, A, fs, n, d);
    const float inv_a0 = 1.0f / d[0];                   /* :160-165 */
    bq->b0 = n[0] * inv_a0;
    bq->b1 = n[1] * inv_a0;
    bq->b2 = n[2] * inv_a0;
    bq->a1 = d[1] * inv_a0;
    bq->a2 = d[2] * inv_a0;
}

/* (int32_t) cast with the firmware's saturating semantics (soft-float
 * __aeabi_f2iz on the RP2040); in range for every valid coefficient. */
static int32_t to_q28(float v)
{
    const float x = v * (float)(1LL << 28);
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

void dspi_compute_coefficients_q28(dspi_eq_param *p, dspi_biquad_q28 *bq, float fs)
{
    if (recipe_is_flat(p) || fs == 0) {
        bq->bypass = 1;
        bq->b0 = 1 << 28;
        bq->b1 = bq->b2 = bq->a1 = bq->a2 = 0;
        return;
    }
    bq->bypass = 0;
    recipe_clamp(p, fs);
    const float A = powf(10.0f, p->gain_db / 40.0f);
    float n[3], d[3];
    cookbook(p, A, fs, n, d);
    bq->b0 = to_q28(n[0] / d[0]);                       /* :168-173 truncating store */
    bq->b1 = to_q28(n[1] / d[0]);
    bq->b2 = to_q28(n[2] / d[0]);
    bq->a1 = to_q28(d[1] / d[0]);
    bq->a2 = to_q28(d[2] / d[0]);
}
