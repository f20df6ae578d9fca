/*
 * fw_stub_impl.c — definitions for the SDK / pico-extras / firmware functions that usb_audio.c and
 * pdm_generator.c reference but that never run when the shims drive process_audio_packet() or the PDM loop
 * (USB stack, S/PDIF / I2S drivers, ADC, flash presets, ...).  TEST INFRASTRUCTURE (oracle/_ref only).
 * Each one aborts with its name if it is ever reached, so a shim that strays off the DSP path fails loudly
 * instead of computing with a fake.  Hardware no-ops that the PDM loop does call (PIO / DMA set-up) return
 * quietly.  This file includes none of the real headers on purpose: the symbols only have to exist.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define UNREACHABLE(name) void name(void) { fprintf(stderr, "oracle/_ref: firmware neighbour %s() reached\n", #name); abort(); }
#define QUIET(name)       long name(void) { return 0; }

/* storage for data symbols owned by translation units that are not linked */
#define DATA(name, bytes) char name[bytes] __attribute__((aligned(64)))

#ifdef FW_STUBS_FOR_CHAIN
UNREACHABLE(adc_init) UNREACHABLE(adc_read) UNREACHABLE(adc_select_input) UNREACHABLE(adc_set_temp_sensor_enabled)
UNREACHABLE(audio_i2s_change_data_pin) UNREACHABLE(audio_i2s_mck_change_pin) UNREACHABLE(audio_i2s_mck_set_enabled)
UNREACHABLE(audio_i2s_mck_setup) UNREACHABLE(audio_i2s_mck_update_frequency) UNREACHABLE(audio_i2s_set_enabled)
UNREACHABLE(audio_i2s_setup) UNREACHABLE(audio_i2s_connect_extra) UNREACHABLE(audio_i2s_change_pins)
UNREACHABLE(audio_new_producer_pool) UNREACHABLE(audio_spdif_change_pin) UNREACHABLE(audio_spdif_connect_extra)
UNREACHABLE(audio_spdif_enable_sync) UNREACHABLE(audio_spdif_get_dma_starvations) UNREACHABLE(audio_spdif_get_dma_starvations_instance)
UNREACHABLE(audio_spdif_reset_dma_starvations) UNREACHABLE(audio_spdif_set_enabled) UNREACHABLE(audio_spdif_set_starvation_monitoring)
UNREACHABLE(audio_spdif_setup) UNREACHABLE(bulk_params_collect) UNREACHABLE(bulk_params_apply) UNREACHABLE(busy_wait_ms)
UNREACHABLE(clock_get_hz) UNREACHABLE(fb_ctrl_stream_stop) UNREACHABLE(fb_ctrl_init) UNREACHABLE(fb_ctrl_reset) UNREACHABLE(fb_ctrl_sof_update)
UNREACHABLE(flash_load_params) UNREACHABLE(flash_save_params) UNREACHABLE(irq_set_priority)
UNREACHABLE(pdm_change_pin) UNREACHABLE(pdm_get_dma_fill_pct) UNREACHABLE(pdm_get_ring_fill_pct) UNREACHABLE(pdm_set_enabled)
UNREACHABLE(preset_get_active) UNREACHABLE(preset_get_directory) UNREACHABLE(preset_get_name) UNREACHABLE(preset_get_saved_master_volume)
UNREACHABLE(reset_usb_boot) UNREACHABLE(spin_lock_blocking) UNREACHABLE(spin_unlock)
UNREACHABLE(usb_current_in_packet_buffer) UNREACHABLE(usb_current_out_packet_buffer) UNREACHABLE(usb_device_init)
UNREACHABLE(usb_device_start) UNREACHABLE(usb_grow_transfer) UNREACHABLE(usb_interface_init) UNREACHABLE(usb_packet_done)
UNREACHABLE(usb_set_default_transfer) UNREACHABLE(usb_start_control_out_transfer)
UNREACHABLE(usb_start_empty_control_in_transfer_null_completion) UNREACHABLE(usb_start_empty_transfer)
UNREACHABLE(usb_start_single_buffer_control_in_transfer) UNREACHABLE(usb_start_tiny_control_in_transfer)
UNREACHABLE(usb_start_transfer) UNREACHABLE(usb_stream_noop_on_chunk) UNREACHABLE(usb_stream_noop_on_packet_complete)
UNREACHABLE(usb_stream_setup_transfer) UNREACHABLE(vreg_get_voltage) UNREACHABLE(usb_get_control_out_endpoint) UNREACHABLE(usb_get_control_in_endpoint)
DATA(audio_device_config, 1024); DATA(boot_device_descriptor, 64); DATA(descriptor_strings, 256);
DATA(fb_ctrl, 256); DATA(feedback_10_14, 8); DATA(nominal_feedback_10_14, 8);
DATA(ms_compat_id_descriptor, 64); DATA(ms_ext_prop_descriptor, 256); DATA(ms_os_string_descriptor, 64);
DATA(output_type_switch_in_progress, 8); DATA(usb_audio_alt_set, 8); DATA(usb_audio_mounted, 8); DATA(usb_audio_packets, 8);
DATA(usb_bitstuff_error_count, 8); DATA(usb_crc_error_count, 8); DATA(usb_data_seq_error_count, 8); DATA(usb_error_count, 8);
DATA(usb_rx_overflow_count, 8); DATA(usb_rx_timeout_count, 8); DATA(usb_descriptor_str_serial, 64);
DATA(usb_control_in, 256); DATA(usb_control_out, 256);
#endif

#ifdef FW_STUBS_FOR_PDM
QUIET(channel_config_set_dreq) QUIET(channel_config_set_read_increment) QUIET(channel_config_set_ring)
QUIET(channel_config_set_transfer_data_size) QUIET(channel_config_set_write_increment) QUIET(dma_channel_abort)
QUIET(dma_channel_configure) QUIET(dma_claim_unused_channel) QUIET(pio_add_program) QUIET(pio_get_dreq)
QUIET(pio_gpio_init) QUIET(pio_sm_init) QUIET(pio_sm_set_clkdiv) QUIET(pio_sm_set_consecutive_pindirs) QUIET(pio_sm_set_enabled)
QUIET(sm_config_set_fifo_join) QUIET(sm_config_set_out_pins) QUIET(sm_config_set_out_shift) QUIET(sm_config_set_wrap)
QUIET(gpio_set_dir) QUIET(gpio_set_function) QUIET(multicore_lockout_victim_init)
unsigned clock_get_hz(int clk) { (void)clk; return 307200000u; }
typedef struct { unsigned ctrl; } dma_channel_config_;
dma_channel_config_ dma_channel_get_default_config(unsigned ch) { dma_channel_config_ c = { ch }; return c; }
typedef struct { unsigned a, b, c, d; } pio_sm_config_;
pio_sm_config_ pio_get_default_sm_config(void) { pio_sm_config_ c = { 0, 0, 0, 0 }; return c; }
#endif
