// wide32.cuh -- int32 kernel family for pairs the s16x2 record cannot hold (placeholder wiring;
// the kernels land in the next milestone).  Until then such pairs are rejected loudly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/b2align.h"

struct b2a_ctx;
namespace b2a {
struct WideState {
    int plan(b2a_ctx*, const std::vector<uint32_t>&, const uint64_t*, const uint64_t*, const b2a_params&, bool);
    int fill(b2a_ctx*, cudaStream_t, uint64_t*);
    int traceback(b2a_ctx*, cudaStream_t, uint64_t*);
    void release() {}
};
} // namespace b2a
