// kernels_tiled.cuh — shared-memory tiled sweep kernels (single-GPU fast path).
// Placeholder plan: disabled until the tiled kernels land; the v1 kernels are the path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_v1.cuh"

namespace lqft {

struct TiledPlan {
    bool enabled = false;
};

inline cudaError_t tiled_prepare(TiledPlan& p, const Geom&, uint32_t, int, bool) {
    p.enabled = false;
    return cudaSuccess;
}
inline void tiled_release(TiledPlan&) {}
inline cudaError_t tiled_sweep(TiledPlan&, const Geom&, int32_t* const*, uint64_t, const uint64_t*,
                               const ChainDev*, const uint32_t*, unsigned long long*, uint32_t, bool,
                               cudaStream_t, uint64_t*) {
    return cudaErrorNotSupported;
}

}  // namespace lqft
