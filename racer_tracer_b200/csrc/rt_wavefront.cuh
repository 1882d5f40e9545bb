// rt_wavefront.cuh — wavefront variant (placeholder until the kernels land).
#pragma once
#include "rt_kernels.cuh"

struct WavefrontState {
    int dummy = 0;
};

inline void wavefront_release(WavefrontState&) {}

inline int wavefront_render(WavefrontState&, int, const KParams&, float*, int, int, size_t, cudaStream_t, int, uint64_t&) {
    return RC_ERR_INVALID;
}
