/* client.c — a plain C11 caller of include/dspi_b200.h, the way firmware-side host code would bind the library
 * (INTEGRATION.md).  Compiled by tests/test_c_client.py with gcc -std=c11 -Wall -Wextra -Werror -pedantic, so the
 * header has to stay valid C.  Without a GPU it exercises the host-side parameter API and expects DSPI_ENODEV from
 * create; with one (argv[1] == "gpu") it runs the reference's 2-channel, 3-band configuration through the engine and
 * compares the result with the same cascade computed in this file from the downloaded coefficients (strict float). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dspi_b200.h"

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d: %s failed (%s)\n", __FILE__, __LINE__, #cond, dspi_last_error()); return 1; } } while (0)

/* dsp_pipeline.c:347-362 / :299-342, strict float */
static void cascade(dspi_biquad_f32 *bq, int nb, float *x, int n)
{
    for (int b = 0; b < nb; b++) {
        dspi_biquad_f32 *q = &bq[b];
        if (q->bypass) continue;
        for (int i = 0; i < n; i++) {
            const float in = x[i];
            if (q->use_svf) {
                const float v3 = in - q->svic2eq;
                const float v1 = q->sva1 * q->svic1eq + q->sva2 * v3;
                const float v2 = q->svic2eq + q->sva2 * q->svic1eq + q->sva3 * v3;
                q->svic1eq = 2.0f * v1 - q->svic1eq;
                q->svic2eq = 2.0f * v2 - q->svic2eq;
                x[i] = q->svm0 * in + q->svm1 * v1 + q->svm2 * v2;
            } else {
                const float out = q->b0 * in + q->s1;
                q->s1 = q->b1 * in - q->a1 * out + q->s2;
                q->s2 = q->b2 * in - q->a2 * out;
                x[i] = out;
            }
        }
    }
}

int main(int argc, char **argv)
{
    const float fs = 48000.0f;
    /* BASELINE config 1: low shelf 100 Hz +4 dB, peaking 1 kHz -3 dB Q 1.4, high shelf 10 kHz +2 dB */
    dspi_eq_param recipe[3] = { { 0, 0, DSPI_FILTER_LOWSHELF, 0, 100.0f, 0.707f, 4.0f }, { 0, 1, DSPI_FILTER_PEAKING, 0, 1000.0f, 1.4f, -3.0f },
                                { 0, 2, DSPI_FILTER_HIGHSHELF, 0, 10000.0f, 0.707f, 2.0f } };
    dspi_biquad_f32 bq[2][DSPI_MAX_BANDS];
    memset(bq, 0, sizeof bq);
    for (int ch = 0; ch < 2; ch++)
        for (int b = 0; b < DSPI_MAX_BANDS; b++) {
            if (b < 3) { dspi_eq_param p = recipe[b]; dspi_compute_coefficients_f32(&p, &bq[ch][b], fs); }
            else { bq[ch][b].bypass = 1; bq[ch][b].b0 = 1.0f; }
        }
    CHECK(bq[0][0].use_svf == 1 && bq[0][1].use_svf == 1 && bq[0][2].use_svf == 0);      /* f < Fs / 7.5 -> SVF (dsp_pipeline.c:88) */
    CHECK(dspi_delay_samples(10.0f, fs, 0) == 480 && dspi_delay_samples(0.0f, fs, 1) == 128);
    CHECK(dspi_delay_samples(1000.0f, fs, 0) == DSPI_CHAIN_MAX_DELAY);
    uint8_t row = 0;
    CHECK(dspi_host_volume(0, &row) == -32768 && row == 60);                              /* the int16 quirk */
    CHECK(dspi_host_volume(-20 * 256, &row) == 0x0ccd && row == 40);
    float lin; int32_t q;
    CHECK(dspi_preamp(0.0f, &lin, &q) == 0 && lin == 1.0f && q == (1 << 28));
    CHECK(dspi_preamp(NAN, &lin, &q) == -1);
    CHECK(dspi_master_volume(-128.0f, &lin, &q) == 0 && lin == 0.0f && q == 0);
    dspi_preset_mute m = { 0, { 0, 0, 0 }, 0, 1.0f };
    dspi_preset_mute_arm(&m, 48000);
    CHECK(m.loading == 1 && m.counter == 512);
    float g = 1.0f;
    for (int p = 0; p < 30; p++) g = dspi_preset_mute_step(&m, 48, 48000);
    CHECK(g == 1.0f && m.loading == 0);
    CHECK(dspi_crc32("123456789", 9) == 0xCBF43926u);

    dspi_eq *e = NULL;
    dspi_eq_desc d = { DSPI_ARITH_F32_STRICT, 2, 10, 0, 0 };
    const int rc = dspi_eq_create(&e, &d);
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
        if (dspi_device_count() == 0) CHECK(rc == DSPI_ENODEV && e == NULL && strlen(dspi_last_error()) > 0);
        else if (rc == DSPI_OK) dspi_eq_destroy(e);
        printf("host api ok\n");
        return 0;
    }
    CHECK(rc == DSPI_OK);
    enum { T = 480 };
    float *x = (float *)dspi_host_alloc(2 * T * sizeof(float));
    float want[2][T];
    CHECK(x != NULL);
    for (int ch = 0; ch < 2; ch++)
        for (int i = 0; i < T; i++) want[ch][i] = x[ch * T + i] = (i == 0) ? 0.5f : 0.25f * sinf(0.13f * (float)(i * (ch + 1)));
    CHECK(dspi_eq_upload_biquads(e, 0, 2, bq) == DSPI_OK);
    CHECK(dspi_eq_process_host(e, x, T) == DSPI_OK);
    for (int ch = 0; ch < 2; ch++) cascade(bq[ch], 10, want[ch], T);
    for (int ch = 0; ch < 2; ch++)
        for (int i = 0; i < T; i++) CHECK(memcmp(&want[ch][i], &x[ch * T + i], 4) == 0);
    dspi_biquad_f32 back[2][DSPI_MAX_BANDS];
    CHECK(dspi_eq_download_biquads(e, 0, 2, back) == DSPI_OK);
    CHECK(memcmp(&back[1][2].s1, &bq[1][2].s1, 8) == 0 && memcmp(&back[0][0].svic1eq, &bq[0][0].svic1eq, 8) == 0);
    dspi_host_free(x);
    CHECK(dspi_eq_destroy(e) == DSPI_OK);
    printf("gpu ok\n");
    return 0;
}
