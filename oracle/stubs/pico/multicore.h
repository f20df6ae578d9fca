/* Stand-in for pico/multicore.h: the lockout handshake is declared by ref_preset_shim.c.  Test infrastructure. */
#pragma once
#include <stdbool.h>
bool multicore_lockout_victim_is_initialized(unsigned core);
void multicore_lockout_start_blocking(void);
void multicore_lockout_end_blocking(void);
unsigned __get_current_exception(void);
