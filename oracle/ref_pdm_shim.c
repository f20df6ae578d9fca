/*
 * ref_pdm_shim.c — runs the reference's OWN delta-sigma loop (pdm_generator.c:204-420, entered through
 * pdm_core1_entry() :691-721) on the host.
 *
 * TEST INFRASTRUCTURE (oracle/_ref only).  pdm_generator.c is #included from where it lies under
 * /root/reference, unmodified.  On the device the loop is paced by the PIO's DMA read pointer
 * (dma_hw->ch[].read_addr) and fed by core 0 through the 256-entry ring (pdm_push_sample).  Here
 * `dma_hw` (oracle/stubs_fw/fw_stub.h) resolves to ref_dma_hw() below, which plays both roles:
 *   - it reports a read pointer exactly TARGET_LEAD (256 words) behind the loop's write pointer, i.e. the
 *     steady state: no underrun recovery, no ring underrun, no overrun (pdm_generator.c:279-317);
 *   - whenever the ring is empty it refills it from the caller's Q28 samples with the reference's own
 *     pdm_push_sample(), and when the input is exhausted it leaves the endless loop with longjmp;
 *   - it copies the words the loop has written to pdm_dma_buffer since the last call to the caller.
 * The modulator state lives in locals of pdm_processing_loop(), so one call = one stream from the
 * restart state (:239-270: integrators, noise shaper and fade-in cleared).  Only fast_rand()'s state is a
 * file-scope static (:62); the caller passes its start value (power-on: 123456789).
 */
#include <setjmp.h>
#include <stdint.h>
#include <string.h>
#include "pico/stdlib.h"
#include "usb_audio.h"
#include "dsp_pipeline.h"

#include "pdm_generator.c"

/* neighbours of translation units that are not linked (only the EQ-worker branch reads them) */
volatile SystemStatusPacket global_status;
MatrixMixer matrix_mixer;
volatile bool bypass_master_eq;
bool channel_bypassed[NUM_CHANNELS];
bool any_delay_active;
int32_t channel_delay_samples[NUM_DELAY_CHANNELS];
Biquad filters[NUM_CHANNELS][MAX_BANDS];
#if PICO_RP2350
float delay_lines[NUM_DELAY_CHANNELS][MAX_DELAY_SAMPLES];
void dsp_process_channel_block(Biquad *restrict b, float *restrict s, uint32_t n, uint8_t ch) { (void)b; (void)s; (void)n; (void)ch; }
#else
int32_t delay_lines[NUM_DELAY_CHANNELS][MAX_DELAY_SAMPLES];
void dsp_process_channel_block(Biquad *restrict b, int32_t *restrict s, uint32_t n, uint8_t ch) { (void)b; (void)s; (void)n; (void)ch; }
#endif
volatile uint32_t pdm_ring_overruns, pdm_ring_underruns, pdm_dma_overruns, pdm_dma_underruns;
pio_hw_t ref_pio_hw[3];
uint32_t time_us_32(void) { static uint32_t t; return t += 3; }
void ref_fw_event(void) { }

static jmp_buf done;
static const int32_t *in_q28;
static uint32_t n_in, in_pos, n_out, cap_idx;
static uint32_t *out_words;
static dma_hw_t hw;
#define WMASK (PDM_DMA_BUFFER_SIZE - 1)

dma_hw_t *ref_dma_hw(void)
{
    while (cap_idx != pdm_stats_write_idx) {                 /* words written since the last call (:380-383) */
        out_words[n_out++] = pdm_dma_buffer[cap_idx];
        cap_idx = (cap_idx + 1) & WMASK;
    }
    if (pdm_head == pdm_tail) {                              /* ring drained: core 0's next packets, or the end */
        if (in_pos == n_in) longjmp(done, 1);
        for (int k = 0; k < 192 && in_pos < n_in; k++) pdm_push_sample(in_q28[in_pos++], false);
    }
    hw.ch[pdm_dma_chan].read_addr = (uint32_t)(uintptr_t)pdm_dma_buffer + 4u * ((pdm_stats_write_idx - 256u) & WMASK);
    return &hw;
}

/* n Q28 samples -> 8 words (256 one-bit decisions, MSB first) each.  Returns the words written; *rng_io is
 * fast_rand()'s state before / after.  ring_overruns etc. must stay 0 (returned through counters[4]). */
uint32_t ref_pdm_run(const int32_t *samples_q28, uint32_t n, uint32_t *rng_io, uint32_t *words_out, uint32_t counters[4])
{
    in_q28 = samples_q28; n_in = n; in_pos = 0; out_words = words_out; n_out = 0;
    if (pdm_dma_chan < 0) pdm_setup_hw(PICO_PDM_PIN);        /* :124-154, PIO/DMA calls are no-ops */
    rng_state = *rng_io;
    cap_idx = pdm_stats_write_idx;
    pdm_head = pdm_tail = 0;
    pdm_ring_overruns = pdm_ring_underruns = pdm_dma_overruns = pdm_dma_underruns = 0;
    core1_mode = CORE1_MODE_PDM;
    pdm_enabled = true;
    if (!setjmp(done)) pdm_core1_entry();
    *rng_io = rng_state;
    if (counters) { counters[0] = pdm_ring_overruns; counters[1] = pdm_ring_underruns; counters[2] = pdm_dma_overruns; counters[3] = pdm_dma_underruns; }
    return n_out;
}
