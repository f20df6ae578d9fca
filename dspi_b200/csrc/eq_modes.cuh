// eq_modes.cuh — band topology codes of the packed coefficient store (safe for runtime compilation).
#pragma once
#include "dspi_common.cuh"

namespace dspi {

constexpr int kMaxBands = 12;      // DSPI_MAX_BANDS / config.h:329

// per-band topology of one channel, 4 bits per band in a 64-bit word (band b -> bits 4b..4b+3)
enum : uint32_t {
    kModeBypass = 0,   // Biquad.bypass                       (dsp_pipeline.c:288)
    kModeTdf2   = 1,   // !use_svf                            (dsp_pipeline.c:347-362)
    kModeSvfLP  = 2,   // use_svf, svf_type == FILTER_LOWPASS (dsp_pipeline.c:299-309)
    kModeSvfHP  = 3,   //                    FILTER_HIGHPASS  (:310-320)
    kModeSvfPK  = 4,   //                    FILTER_PEAKING   (:321-331)
    kModeSvfSH  = 5    // default: shelves, general mix       (:332-342)
};

}  // namespace dspi
