/* orc_spdif.c — CPU restatement of the S/PDIF (IEC 60958) subframe encoder that consumes the chain's
 * 24-bit words.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / reference legs of bench.py may call it.
 *
 * Reference (firmware/pico-extras/src/rp2_common/pico_audio_spdif_multi/):
 *   lookup table          audio_spdif.c:141-153
 *   spdif_update_subframe include/pico/audio_spdif/sample_encoding.h:27-50
 *   preambles, channel status, initial h      audio_spdif.c:73-114  (init_spdif_buffer)
 *   block-position fix-up at DMA start        audio_spdif.c:372-388
 * Pinned: orc_spdif_update_subframe against the reference's own inline function compiled from its
 * header (oracle/_ref/libdspi_ref_spdif.so, tests/test_spdif_cpu.py).  The table fill and the
 * preamble / channel-status stamping need the SDK to compile and are restated only ("parity
 * unpinned" for those ten lines; the tests check them against IEC 60958 properties instead). */
#include <stdint.h>
#include "dspi_oracle.h"

#define PREAMBLE_X 0xC9u   /* 0b11001001, audio_spdif.c:77 */
#define PREAMBLE_Y 0x69u   /* 0b01101001, :78 */
#define PREAMBLE_Z 0x39u   /* 0b00111001, :79 */

/* audio_spdif.c:141-153: bit j of the byte -> cell j (2 bits) of a biphase-mark word whose cells all
 * start with a transition (0x5555); a set data bit adds the mid-cell transition; bit 16 = byte parity */
void orc_spdif_lookup_init(uint32_t table[256])
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t v = 0x5555, p = 0;
        for (uint32_t j = 0; j < 8; j++) {
            if (i & (1u << j)) {
                p ^= 1;
                v |= 2u << (j * 2);
            }
        }
        table[i] = v | (p << 16);
    }
}

/* sample_encoding.h:27-50 */
void orc_spdif_update_subframe(const uint32_t table[256], uint32_t *l, uint32_t *h, int32_t sample)
{
    const uint32_t s0 = table[(uint8_t)sample];
    const uint32_t s1 = table[(uint8_t)((uint32_t)sample >> 8)];
    const uint32_t s2 = table[(uint8_t)((uint32_t)sample >> 16)];
    *l = (*l & 0xffu) | ((uint32_t)(uint16_t)s0 << 8) | (s1 << 24);
    const uint32_t ph = *h >> 24;
    const uint32_t hh = ((uint32_t)(uint16_t)s1 >> 8) | ((uint32_t)(uint16_t)s2 << 8);
    uint32_t p = (s0 >> 16) ^ (s1 >> 16) ^ (s2 >> 16);
    p ^= (((ph & 0x2au) * 0x2au) >> 6) & 1u;
    *h = hh | ((ph & 0x7fu) << 24) | (p << 31);
}

/* audio_spdif.c:91-94 */
static uint32_t cstatus_bit(const uint8_t cs[5], uint32_t pos)
{
    if (pos >= 40) return 0;
    return (cs[pos / 8] >> (pos % 8)) & 1u;
}

/* One stereo stream: frame n sits at block position (pos0 + n) % 192.  Left subframe: preamble Z at
 * position 0, X elsewhere; right: Y (init_spdif_buffer :104-111, fix-up :372-388); both carry the
 * channel-status bit of the position in bit 29 of h (initial h = 0x55000000 | c << 29, :106/:109). */
void orc_spdif_encode(const uint32_t table[256], const int32_t *words, uint32_t frames, uint32_t pos0, const uint8_t cs[5], uint32_t *out)
{
    for (uint32_t n = 0; n < frames; n++) {
        const uint32_t pos = (pos0 + n) % 192u;
        const uint32_t c = cstatus_bit(cs, pos);
        for (uint32_t ch = 0; ch < 2; ch++) {
            uint32_t l = ch ? PREAMBLE_Y : (pos == 0 ? PREAMBLE_Z : PREAMBLE_X);
            uint32_t h = 0x55000000u | (c << 29);
            orc_spdif_update_subframe(table, &l, &h, words[2 * n + ch]);
            out[(2 * n + ch) * 2 + 0] = l;
            out[(2 * n + ch) * 2 + 1] = h;
        }
    }
}
