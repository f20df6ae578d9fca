/* Stand-in for the Pico SDK's pico/audio.h, ONLY so that the reference header
 * pico_audio_spdif_multi/include/pico/audio_spdif/sample_encoding.h compiles unmodified on the host
 * (oracle/ref_spdif_shim.c).  It declares the two opaque types that header's prototypes mention and
 * the one SDK macro its inline function uses.  Test infrastructure; nothing of the SDK is restated. */
#ifndef ORC_STUB_PICO_AUDIO_H
#define ORC_STUB_PICO_AUDIO_H
#include <stdint.h>
#include <stdbool.h>
typedef struct audio_connection audio_connection_t;
typedef struct audio_buffer audio_buffer_t;
/* pico/platform: a plain 32-bit multiply the compiler may not strength-reduce */
#define __mul_instruction(a, b) ((uint32_t)(a) * (uint32_t)(b))
#endif
