/* Stand-in for hardware/sync.h (interrupt masking and barriers are no-ops on the host).  Test infrastructure. */
#pragma once
#include <stdint.h>
static inline uint32_t save_and_disable_interrupts(void){return 0;}
static inline void restore_interrupts(uint32_t s){(void)s;}
static inline void __dmb(void){}
static inline void __sev(void){}
static inline void __wfe(void){}
