/* ref_spdif_shim.c — calls into the reference's own spdif_update_subframe()
 * (pico_audio_spdif_multi/include/pico/audio_spdif/sample_encoding.h:27-50), compiled unmodified from
 * where it lies; pico/audio.h is the stub under oracle/stubs.  The 256-entry table that function reads
 * is filled by audio_spdif_setup() (audio_spdif.c:141-153), which needs the SDK: the caller passes the
 * table in (tests fill it from the restatement, orc_spdif_lookup_init).  TEST INFRASTRUCTURE. */
#include "pico/audio_spdif/sample_encoding.h"

uint32_t spdif_lookup[256];

void ref_spdif_set_lookup(const uint32_t *table)
{
    for (int i = 0; i < 256; i++) spdif_lookup[i] = table[i];
}

/* lh[2] = {l, h} in and out */
void ref_spdif_update_subframe(uint32_t *lh, int32_t sample)
{
    spdif_subframe_t sf = { lh[0], lh[1] };
    spdif_update_subframe(&sf, sample);
    lh[0] = sf.l;
    lh[1] = sf.h;
}

/* stereo_to_spdif_producer_give_s32 -> converting_copy<Stereo<FmtSPDIF>, Stereo<FmtS32>>::copy
 * (sample_encoding.cpp:42-50): every subframe of the buffer updated in place with the next word */
void ref_spdif_copy_s32(uint32_t *subframes, const int32_t *src, uint32_t sample_count)
{
    for (uint32_t i = 0; i < sample_count * 2; i++) ref_spdif_update_subframe(subframes + 2 * i, src[i]);
}
