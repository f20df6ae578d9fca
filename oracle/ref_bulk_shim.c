/* ref_bulk_shim.c — runs the reference's own bulk_params_apply() / bulk_params_collect()
 * (firmware/DSPi/bulk_params.c, compiled unmodified from where it lies, once per platform) and copies
 * the globals they touch into / out of a dspi_bulk_state.  This file only DEFINES those globals (the
 * firmware defines them in usb_audio.c, dsp_pipeline.c and main.c, which need the Pico SDK) and moves
 * values; it contains no parameter logic.  TEST INFRASTRUCTURE. */
#include <string.h>
#include "bulk_params.h"
#include "config.h"
#include "dsp_pipeline.h"
#include "usb_audio.h"
#include "crossfeed.h"
#include "leveller.h"
#include "dspi_b200.h"

volatile float global_preamp_db[NUM_INPUT_CHANNELS];
volatile int32_t global_preamp_mul[NUM_INPUT_CHANNELS];
volatile float global_preamp_linear[NUM_INPUT_CHANNELS];
volatile float master_volume_db;
volatile float master_volume_linear;
volatile int32_t master_volume_q15;
volatile float channel_gain_db[3];
volatile int32_t channel_gain_mul[3];
volatile float channel_gain_linear[3];
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

int ref_bulk_platform(void) { return PICO_RP2350 ? DSPI_PLATFORM_RP2350 : DSPI_PLATFORM_RP2040; }
size_t ref_bulk_wire_size(void) { return sizeof(WireBulkParams); }

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
    for (int i = 0; i < 3; i++) {
        channel_gain_db[i] = st->legacy_gain_db[i]; channel_gain_linear[i] = st->legacy_gain_linear[i];
        channel_gain_mul[i] = st->legacy_gain_mul[i]; channel_mute[i] = st->legacy_mute[i];
    }
    for (int i = 0; i < NUM_CHANNELS; i++) channel_delays_ms[i] = st->channel_delays_ms[i];
    for (int in = 0; in < NUM_INPUT_CHANNELS; in++)
        for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&matrix_mixer.crosspoints[in][o], &st->crosspoints[in][o], sizeof(MatrixCrosspoint));
    for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&matrix_mixer.outputs[o], &st->outputs[o], sizeof(OutputChannel));
    for (int ch = 0; ch < NUM_CHANNELS; ch++)
        for (int b = 0; b < MAX_BANDS; b++) memcpy(&filter_recipes[ch][b], &st->recipes[ch][b], sizeof(EqParamPacket));
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
    for (int i = 0; i < 3; i++) {
        st->legacy_gain_db[i] = channel_gain_db[i]; st->legacy_gain_linear[i] = channel_gain_linear[i];
        st->legacy_gain_mul[i] = channel_gain_mul[i]; st->legacy_mute[i] = channel_mute[i];
    }
    for (int i = 0; i < NUM_CHANNELS; i++) st->channel_delays_ms[i] = channel_delays_ms[i];
    for (int in = 0; in < NUM_INPUT_CHANNELS; in++)
        for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&st->crosspoints[in][o], &matrix_mixer.crosspoints[in][o], sizeof(MatrixCrosspoint));
    for (int o = 0; o < NUM_OUTPUT_CHANNELS; o++) memcpy(&st->outputs[o], &matrix_mixer.outputs[o], sizeof(OutputChannel));
    for (int ch = 0; ch < NUM_CHANNELS; ch++)
        for (int b = 0; b < MAX_BANDS; b++) memcpy(&st->recipes[ch][b], &filter_recipes[ch][b], sizeof(EqParamPacket));
}

/* `st` in: the device state before the packet arrives; out: after bulk_params_apply() */
int ref_bulk_apply(const void *wire, dspi_bulk_state *st)
{
    state_to_globals(st);
    const int rc = bulk_params_apply((const WireBulkParams *)wire, false);
    globals_to_state(st);
    return rc;
}

void ref_bulk_collect(const dspi_bulk_state *st, void *wire)
{
    state_to_globals(st);
    memset(output_pins, 0, sizeof(output_pins));
    memset(channel_names, 0, sizeof(channel_names));
    memset(output_types, 0, sizeof(output_types));
    i2s_bck_pin = i2s_mck_pin = 0; i2s_mck_enabled = false; i2s_mck_multiplier = 128;
    bulk_params_collect((WireBulkParams *)wire);
}
