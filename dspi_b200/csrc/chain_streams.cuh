// chain_streams.cuh — stream/event plumbing shared by the float and the Q28 chain engines.
//
// One process call is cut into packet slices and the three stages of a slice (front, outputs, PDM
// modulator) run on three streams chained by events: stage k of slice i+1 overlaps stage k+1 of slice i.
// Each stage is a set of serial recurrences with too few warps to fill the machine on its own (the
// modulator: one warp per 32 instances), so running them side by side is what fills it.  All state
// lives in HBM between slices and every slice touches its own frames of the intermediate buffers, so
// neither the slicing nor the overlap changes a bit.
#pragma once
#include <cuda_runtime.h>

namespace dspi {

struct ChainStreams {
    static constexpr int kMaxSlices = 8;
    cudaStream_t s_front = nullptr, s_out = nullptr, s_pdm = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_done = nullptr, ev_aux = nullptr, ev_front[kMaxSlices] = {}, ev_out[kMaxSlices] = {};

    cudaError_t create()
    {
        int lo = 0, hi = 0;                                   // numerically lower = higher priority
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        // the modulator is the longest serial chain: its few CTAs are placed first whenever an SM frees a
        // slot; then the front; the many output CTAs fill what is left
        const int mid = hi < lo ? hi + 1 : lo;
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s_pdm, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s_front, cudaStreamNonBlocking, mid);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s_out, cudaStreamNonBlocking, lo);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_begin, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_aux, cudaEventDisableTiming);
        for (int i = 0; i < kMaxSlices && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&ev_front[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming);
        }
        return e;
    }

    void destroy()
    {
        for (cudaStream_t *s : { &s_front, &s_out, &s_pdm })
            if (*s) { cudaStreamSynchronize(*s); cudaStreamDestroy(*s); *s = nullptr; }
        for (cudaEvent_t *ev : { &ev_begin, &ev_done, &ev_aux })
            if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
        for (int i = 0; i < kMaxSlices; i++) {
            if (ev_front[i]) { cudaEventDestroy(ev_front[i]); ev_front[i] = nullptr; }
            if (ev_out[i]) { cudaEventDestroy(ev_out[i]); ev_out[i] = nullptr; }
        }
    }
};

}  // namespace dspi
