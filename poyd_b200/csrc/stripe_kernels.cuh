// Register-resident stripe kernels (placeholder until the first generic-path GPU validation lands).
#pragma once
#include "cells.cuh"

namespace poyb200 {
constexpr uint32_t KLASS_GENERIC = 0;
static inline bool stripe_choose(Task &, bool, int) { return false; }
static inline cudaError_t stripe_launch(uint32_t, bool, bool, const Task *, int, DevCM, const uint8_t *, uint8_t *, int *, int,
                                        cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace poyb200
