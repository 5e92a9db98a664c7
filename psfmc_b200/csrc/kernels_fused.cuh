// Fused single-kernel shared-memory path for frames up to 128^2 (float32).
// Placeholder until the fused kernel lands: the engine uses the staged path.
#pragma once
#include "pipeline.cuh"

namespace psfmc {

template <typename T>
inline bool fused_path_available(const StagedPlan &, const Program &) {
  return false;
}

template <typename T>
inline int fused_prepare_device(const StagedPlan &) {
  return 0;
}

template <typename T>
inline int launch_fused_lnlike(const StagedPlan &, const StagedBuffers<T> &, int, int,
                               const double *, long long, long long, double *,
                               cudaStream_t) {
  return 0;
}

}  // namespace psfmc
