/* ref_preset_shim.c — runs the reference's own preset_save() / preset_load() (firmware/DSPi/flash_storage.c,
 * compiled unmodified from where it lies, once per platform) against a RAM image of the flash and moves the
 * globals they touch into / out of a dspi_bulk_state.  This file DEFINES the globals and the hardware / other-
 * module functions flash_storage.c links against (flash erase/program on the RAM image, everything else a
 * no-op); it contains no preset logic.  TEST INFRASTRUCTURE. */
#include <string.h>
#include "flash_storage.h"
#include "config.h"
#include "dsp_pipeline.h"
#include "usb_audio.h"
#include "crossfeed.h"
#include "leveller.h"
#include "pdm_generator.h"
#include "usb_feedback_controller.h"
#include "hardware/flash.h"
#include "dspi_b200.h"

uint8_t ref_flash_image[PICO_FLASH_SIZE_BYTES];

/* ---- globals of usb_audio.c / dsp_pipeline.c / main.c that flash_storage.c reads and writes ---- */
volatile float global_preamp_db[NUM_INPUT_CHANNELS];
volatile int32_t global_preamp_mul[NUM_INPUT_CHANNELS];
volatile float global_preamp_linear[NUM_INPUT_CHANNELS];
volatile float master_volume_db;
volatile float master_volume_linear;
volatile int32_t master_volume_q15;
volatile float channel_gain_db[3];
volatile int32_t channel_gain_mul[3];
volatile bool channel_mute[3];
volatile bool loudness_enabled;
volatile float loudness_ref_spl;
volatile float loudness_intensity_pct;
volatile bool loudness_recompute_pending;
volatile CrossfeedConfig crossfeed_config;
volatile bool crossfeed_update_pending;
volatile LevellerConfig leveller_config;
volatile bool leveller_update_pending;
volatile bool leveller_reset_pending;
MatrixMixer matrix_mixer;
uint8_t output_pins[NUM_PIN_OUTPUTS];
EqParamPacket filter_recipes[NUM_CHANNELS][MAX_BANDS];
float channel_delays_ms[NUM_CHANNELS];
volatile bool bypass_master_eq;
char channel_names[NUM_CHANNELS][PRESET_NAME_LEN];
uint8_t output_types[NUM_SPDIF_INSTANCES];
uint8_t i2s_bck_pin, i2s_mck_pin;
bool i2s_mck_enabled;
uint16_t i2s_mck_multiplier;
volatile AudioState audio_state;
volatile Core1Mode core1_mode;
#if PICO_RP2350
float delay_lines[NUM_DELAY_CHANNELS][MAX_DELAY_SAMPLES];
#else
int32_t delay_lines[NUM_DELAY_CHANNELS][MAX_DELAY_SAMPLES];
#endif
volatile uint32_t feedback_10_14, nominal_feedback_10_14;
usb_feedback_ctrl_t fb_ctrl;

/* ---- other modules / hardware: no-ops, except the flash which is the RAM image ---- */
void dspi_flash_range_erase(uint32_t off, size_t n) { memset(ref_flash_image + off, 0xFF, n); }
void dspi_flash_range_program(uint32_t off, const uint8_t *d, size_t n) { memcpy(ref_flash_image + off, d, n); }
bool multicore_lockout_victim_is_initialized(unsigned core) { (void)core; return false; }
void multicore_lockout_start_blocking(void) {}
void multicore_lockout_end_blocking(void) {}
unsigned __get_current_exception(void) { return 0; }
void fb_ctrl_reset(usb_feedback_ctrl_t *c, uint32_t nominal) { (void)c; (void)nominal; }
void dsp_recalculate_all_filters(float fs) { (void)fs; }
void dsp_update_delay_samples(float fs) { (void)fs; }
void dsp_init_default_filters(void) {}
void get_default_channel_name(int ch, char *buf) { (void)ch; buf[0] = 0; }
Core1Mode derive_core1_mode(void) { return core1_mode; }
void pdm_set_enabled(bool on) { (void)on; }

int ref_preset_platform(void) { return PICO_RP2350 ? DSPI_PLATFORM_RP2350 : DSPI_PLATFORM_RP2040; }

static void state_to_globals(const dspi_bulk_state *st)
{
    for (int i = 0; i < NUM_INPUT_CHANNELS; i++) {
        global_preamp_db[i] = st->preamp_db[i]; global_preamp_linear[i] = st->preamp_linear[i]; global_preamp_mul[i] = st->preamp_q28[i];
    }
    master_volume_db = st->master_volume_db; master_volume_linear = st->master_volume_linear; master_volume_q15 = st->master_volume_q15;
    bypass_master_eq = st->bypass_master_eq; loudness_enabled = st->loudness_enabled;
    loudness_ref_spl = st->loudness_ref_spl; loudness_intensity_pct = st->loudness_intensity_pct;
    crossfeed_config.enabled = st->crossfeed.enabled; crossfeed_config.itd_enabled = st->crossfeed.itd_enabled;
    crossfeed_config.preset = st->crossfeed.preset; crossfeed_config.custom_fc = st->crossfeed.custom_fc;
    crossfeed_config.custom_feed_db = st->crossfeed.custom_feed_db;
    leveller_config.enabled = st->leveller.enabled; leveller_config.amount = st->leveller.amount; leveller_config.speed = st->leveller.speed;
    leveller_config.max_gain_db = st->leveller.max_gain_db; leveller_config.lookahead = st->leveller.lookahead;
    leveller_config.gate_threshold_db = st->leveller.gate_threshold_db;
    for (int i = 0; i < 3; i++) { channel_gain_db[i] = st->legacy_gain_db[i]; channel_gain_mul[i] = st->legacy_gain_mul[i]; channel_mute[i] = st->legacy_mute[i]; }
    for (int i = 0; i < NUM_CHANNELS; i++) channel_delays_ms[i] = st->channel_delays_ms[i];
    for (int in = 0; in < NUM_INPUT_CHANNELS; in++)
        for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&matrix_mixer.crosspoints[in][o], &st->crosspoints[in][o], sizeof(MatrixCrosspoint));
    for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&matrix_mixer.outputs[o], &st->outputs[o], sizeof(OutputChannel));
    for (int ch = 0; ch < NUM_CHANNELS; ch++)
        for (int b = 0; b < MAX_BANDS; b++) memcpy(&filter_recipes[ch][b], &st->recipes[ch][b], sizeof(EqParamPacket));
    memset(output_pins, 0, sizeof(output_pins));
    memset(channel_names, 0, sizeof(channel_names));
    memset(output_types, 0, sizeof(output_types));
    i2s_bck_pin = i2s_mck_pin = 0; i2s_mck_enabled = false; i2s_mck_multiplier = 128;
}

static void globals_to_state(dspi_bulk_state *st)
{
    for (int i = 0; i < NUM_INPUT_CHANNELS; i++) {
        st->preamp_db[i] = global_preamp_db[i]; st->preamp_linear[i] = global_preamp_linear[i]; st->preamp_q28[i] = global_preamp_mul[i];
    }
    st->master_volume_db = master_volume_db; st->master_volume_linear = master_volume_linear; st->master_volume_q15 = master_volume_q15;
    st->bypass_master_eq = bypass_master_eq; st->loudness_enabled = loudness_enabled;
    st->loudness_ref_spl = loudness_ref_spl; st->loudness_intensity_pct = loudness_intensity_pct;
    st->crossfeed.enabled = crossfeed_config.enabled; st->crossfeed.itd_enabled = crossfeed_config.itd_enabled;
    st->crossfeed.preset = crossfeed_config.preset; st->crossfeed.custom_fc = crossfeed_config.custom_fc;
    st->crossfeed.custom_feed_db = crossfeed_config.custom_feed_db;
    st->leveller.enabled = leveller_config.enabled; st->leveller.amount = leveller_config.amount; st->leveller.speed = leveller_config.speed;
    st->leveller.max_gain_db = leveller_config.max_gain_db; st->leveller.lookahead = leveller_config.lookahead;
    st->leveller.gate_threshold_db = leveller_config.gate_threshold_db;
    for (int i = 0; i < 3; i++) { st->legacy_gain_db[i] = channel_gain_db[i]; st->legacy_gain_mul[i] = channel_gain_mul[i]; st->legacy_mute[i] = channel_mute[i]; }
    for (int i = 0; i < NUM_CHANNELS; i++) st->channel_delays_ms[i] = channel_delays_ms[i];
    for (int in = 0; in < NUM_INPUT_CHANNELS; in++)
        for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&st->crosspoints[in][o], &matrix_mixer.crosspoints[in][o], sizeof(MatrixCrosspoint));
    for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&st->outputs[o], &matrix_mixer.outputs[o], sizeof(OutputChannel));
    for (int ch = 0; ch < NUM_CHANNELS; ch++)
        for (int b = 0; b < MAX_BANDS; b++) memcpy(&st->recipes[ch][b], &filter_recipes[ch][b], sizeof(EqParamPacket));
}

static uint8_t *slot_sector(uint8_t slot)
{
    return ref_flash_image + (PICO_FLASH_SIZE_BYTES - 12u * FLASH_SECTOR_SIZE) + (1u + slot) * FLASH_SECTOR_SIZE;   /* flash_storage.c:52-57 */
}

/* preset_save(slot) with the device in state `st`; copies the written sector (4096 bytes) to `sector_out` */
int ref_preset_save(const dspi_bulk_state *st, uint8_t slot, uint8_t *sector_out)
{
    state_to_globals(st);
    audio_state.freq = 48000;
    const int rc = preset_save(slot);
    memcpy(sector_out, slot_sector(slot), FLASH_SECTOR_SIZE);
    return rc;
}

/* preset_load(slot) of the image `bytes`, directory master-volume mode / value as given; `st` in: state before, out: after */
int ref_preset_load(const uint8_t *bytes, size_t len, uint8_t slot, uint8_t master_volume_mode, float dir_master_volume_db, dspi_bulk_state *st)
{
    state_to_globals(st);
    audio_state.freq = 48000;
    preset_save(slot);                                        /* marks the slot occupied in the directory */
    memset(slot_sector(slot), 0xFF, FLASH_SECTOR_SIZE);
    memcpy(slot_sector(slot), bytes, len);                    /* the image under test */
    preset_set_master_volume_mode(master_volume_mode);
    master_volume_db = dir_master_volume_db;
    preset_save_master_volume();                              /* directory's independent master volume */
    state_to_globals(st);
    const int rc = preset_load(slot);
    globals_to_state(st);
    return rc;
}
