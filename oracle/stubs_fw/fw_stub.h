/* fw_stub.h — stand-ins for the Pico SDK headers that usb_audio.c / pdm_generator.c include.
 *
 * TEST INFRASTRUCTURE (oracle/_ref builds only).  The reference's own sources are compiled
 * unmodified from /root/reference; the pico-extras headers they include (usb_device.h, audio.h,
 * audio_spdif.h, audio_i2s_multi.h, usb_common.h, buffer.h) are taken from the reference tree too.
 * Only the SDK *platform* layer (registers, clocks, PIO/DMA handles, section attributes) is replaced
 * here by the minimum that lets the translation units compile on x86-64.  None of it carries
 * arithmetic: the DSP path never reads a value produced by these definitions, except
 *   - time_us_32/64 (CPU-load metering, gap detection: not part of any compared output) and
 *   - dma_hw->ch[].read_addr, which ref_pdm_shim.c turns into the pacing hook of the PDM loop. */
#ifndef DSPI_FW_STUB_H
#define DSPI_FW_STUB_H
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <string.h>
#include <assert.h>

typedef unsigned int uint;
typedef uint64_t absolute_time_t;

#define __packed                 __attribute__((packed))
#define __aligned(x)             __attribute__((aligned(x)))
#define __unused                 __attribute__((unused))
#define __not_in_flash(group)
#define __not_in_flash_func(f)   f
#define __time_critical_func(f)  f
#define __no_inline_not_in_flash_func(f) f
#define __scratch_x(group)
#define __scratch_y(group)
#define __force_inline           inline __attribute__((always_inline))
#ifndef count_of
#define count_of(a)              (sizeof(a) / sizeof((a)[0]))
#endif
#ifndef MIN
#define MIN(a, b)                ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b)                ((a) > (b) ? (a) : (b))
#endif
#define tight_loop_contents()    ((void)0)
#define hard_assert              assert
#define invalid_params_if(x, t)  ((void)0)
#define valid_params_if(x, t)    ((void)0)
#define panic(...)               assert(0)
#define PICO_OK                  0

/* barriers / events: single-threaded on the host */
void ref_fw_event(void);              /* defined by the shim; __wfe() ends up here */
#define __dmb()                  __atomic_signal_fence(__ATOMIC_SEQ_CST)
#define __dsb()                  __atomic_signal_fence(__ATOMIC_SEQ_CST)
#define __sev()                  ((void)0)
#define __wfe()                  ref_fw_event()
#define __wfi()                  ref_fw_event()
#define __breakpoint()           ((void)0)

/* hardware/sync.h */
typedef volatile uint32_t spin_lock_t;
uint32_t save_and_disable_interrupts(void);
void restore_interrupts(uint32_t status);
spin_lock_t *spin_lock_init(uint lock_num);
spin_lock_t *spin_lock_instance(uint lock_num);
int  spin_lock_claim_unused(bool required);
uint32_t spin_lock_blocking(spin_lock_t *lock);
void spin_unlock(spin_lock_t *lock, uint32_t saved_irq);

/* hardware/timer.h, pico/time.h */
uint32_t time_us_32(void);
uint64_t time_us_64(void);
void sleep_ms(uint32_t ms);
void sleep_us(uint64_t us);
void busy_wait_us(uint64_t us);
void busy_wait_us_32(uint32_t us);
void busy_wait_ms(uint32_t ms);
absolute_time_t get_absolute_time(void);
uint32_t to_ms_since_boot(absolute_time_t t);

/* hardware/irq.h */
typedef void (*irq_handler_t)(void);
#define DMA_IRQ_0 11
#define DMA_IRQ_1 12
#define PIO0_IRQ_0 7
#define PIO1_IRQ_0 9
#define USBCTRL_IRQ 5
#define PICO_HIGHEST_IRQ_PRIORITY 0x00
#define PICO_DEFAULT_IRQ_PRIORITY 0x80
#define PICO_LOWEST_IRQ_PRIORITY  0xff
void irq_set_priority(uint num, uint8_t prio);
uint irq_get_priority(uint num);
void irq_set_enabled(uint num, bool enabled);
void irq_set_exclusive_handler(uint num, irq_handler_t h);
void irq_add_shared_handler(uint num, irq_handler_t h, uint8_t order);

/* hardware/pio.h */
typedef struct { volatile uint32_t txf[4]; volatile uint32_t rxf[4]; volatile uint32_t fdebug, fstat, ctrl; } pio_hw_t;
typedef pio_hw_t *PIO;
extern pio_hw_t ref_pio_hw[3];
#define pio0 (&ref_pio_hw[0])
#define pio1 (&ref_pio_hw[1])
#define pio2 (&ref_pio_hw[2])
typedef struct pio_program { const uint16_t *instructions; uint8_t length; int8_t origin; uint8_t pio_version; } pio_program_t;
typedef struct { uint32_t clkdiv, execctrl, shiftctrl, pinctrl; } pio_sm_config;
enum pio_fifo_join { PIO_FIFO_JOIN_NONE = 0, PIO_FIFO_JOIN_TX = 1, PIO_FIFO_JOIN_RX = 2 };
uint pio_add_program(PIO pio, const pio_program_t *p);
pio_sm_config pio_get_default_sm_config(void);
void sm_config_set_wrap(pio_sm_config *c, uint a, uint b);
void sm_config_set_out_pins(pio_sm_config *c, uint base, uint n);
void sm_config_set_out_shift(pio_sm_config *c, bool r, bool a, uint t);
void sm_config_set_fifo_join(pio_sm_config *c, enum pio_fifo_join j);
void pio_gpio_init(PIO pio, uint pin);
int  pio_sm_set_consecutive_pindirs(PIO pio, uint sm, uint base, uint n, bool out);
int  pio_sm_init(PIO pio, uint sm, uint offset, const pio_sm_config *c);
void pio_sm_set_enabled(PIO pio, uint sm, bool en);
void pio_sm_set_clkdiv(PIO pio, uint sm, float div);
uint pio_get_dreq(PIO pio, uint sm, bool tx);

/* hardware/dma.h */
typedef struct { volatile uint32_t read_addr, write_addr, transfer_count, ctrl_trig; } dma_channel_hw_t;
typedef struct { dma_channel_hw_t ch[16]; } dma_hw_t;
dma_hw_t *ref_dma_hw(void);           /* defined by the shim: the PDM pacing hook */
#define dma_hw (ref_dma_hw())
typedef struct { uint32_t ctrl; } dma_channel_config;
enum dma_channel_transfer_size { DMA_SIZE_8 = 0, DMA_SIZE_16 = 1, DMA_SIZE_32 = 2 };
int  dma_claim_unused_channel(bool required);
dma_channel_config dma_channel_get_default_config(uint ch);
void channel_config_set_transfer_data_size(dma_channel_config *c, enum dma_channel_transfer_size s);
void channel_config_set_read_increment(dma_channel_config *c, bool incr);
void channel_config_set_write_increment(dma_channel_config *c, bool incr);
void channel_config_set_dreq(dma_channel_config *c, uint dreq);
void channel_config_set_ring(dma_channel_config *c, bool write, uint bits);
void dma_channel_configure(uint ch, const dma_channel_config *c, volatile void *wr, const volatile void *rd, uint32_t n, bool trigger);
void dma_channel_abort(uint ch);

/* hardware/clocks.h */
enum clock_index { clk_gpout0 = 0, clk_ref, clk_sys, clk_peri, clk_usb, clk_adc };
uint32_t clock_get_hz(enum clock_index clk);

/* hardware/gpio.h */
enum gpio_function { GPIO_FUNC_NULL = 0x1f, GPIO_FUNC_PIO0 = 6, GPIO_FUNC_PIO1 = 7, GPIO_FUNC_SIO = 5 };
#define GPIO_IN  false
#define GPIO_OUT true
void gpio_set_function(uint pin, enum gpio_function fn);
void gpio_set_dir(uint pin, bool out);
void gpio_init(uint pin);
void gpio_put(uint pin, bool v);
bool gpio_get(uint pin);
void gpio_disable_pulls(uint pin);

/* hardware/adc.h, hardware/vreg.h */
#define NUM_ADC_CHANNELS 5
void adc_select_input(uint input);
uint16_t adc_read(void);
enum vreg_voltage { VREG_VOLTAGE_0_55 = 0, VREG_VOLTAGE_0_85 = 6, VREG_VOLTAGE_1_10 = 11, VREG_VOLTAGE_1_15 = 12, VREG_VOLTAGE_1_30 = 15, VREG_VOLTAGE_3_30 = 31 };
enum vreg_voltage vreg_get_voltage(void);

/* pico/bootrom.h, pico/multicore.h, hardware/watchdog.h */
void reset_usb_boot(uint32_t gpio_mask, uint32_t disable_mask);
void rom_reset_usb_boot(uint32_t gpio_mask, uint32_t disable_mask);
void multicore_launch_core1(void (*entry)(void));
void multicore_reset_core1(void);
uint get_core_num(void);
void watchdog_update(void);
void watchdog_reboot(uint32_t pc, uint32_t sp, uint32_t delay_ms);

/* usb_common.h wants the controller's endpoint count */
#define USB_NUM_ENDPOINTS 16
#endif
