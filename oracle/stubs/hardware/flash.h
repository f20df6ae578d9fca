/* Stand-in for the Pico SDK's hardware/flash.h so that the reference's flash_storage.c compiles unmodified on
 * the host (oracle/ref_preset_shim.c): the "flash" is a RAM image and XIP reads go to it.  Test infrastructure. */
#ifndef ORC_STUB_HARDWARE_FLASH_H
#define ORC_STUB_HARDWARE_FLASH_H
#include <stdint.h>
#include <stddef.h>
#define FLASH_SECTOR_SIZE 4096u
#define FLASH_PAGE_SIZE 256u
#ifndef PICO_FLASH_SIZE_BYTES
#define PICO_FLASH_SIZE_BYTES (2u * 1024u * 1024u)
#endif
extern uint8_t ref_flash_image[];
#define XIP_BASE ((uintptr_t)ref_flash_image)
#endif
