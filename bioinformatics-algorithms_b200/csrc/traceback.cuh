// traceback.cuh -- one thread per pair walks the short16 record (b2a_format.h walkers).
//
// Replaces hw2.cpp:158-188 (NW) / hw2.cpp:235-263 (SW) and overlapLongestExactMatch (hw2.cpp:267-278)
// for every pair of the batch.  The walk is a dependent chain of 16-byte chunk loads (one per row
// change), so the kernel is latency/HBM bound; parallelism comes from the batch (one thread per pair,
// the two halves of a pair-pair in adjacent threads so they share sectors).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b2a_format.h"
#include "short16_fill.cuh"

namespace b2a {

struct TbArgs {
    const uint8_t*  pat;
    const uint8_t*  txt;
    const uint64_t* pat_off;
    const uint64_t* txt_off;
    const PPDesc*   pps;
    const uint64_t* code_off;
    const Chunk*    codes;
    const uint32_t* rowbest;
    PairResult*     results;    // indexed by pair
    uint32_t*       ops;        // packed 2-bit ops, may be null
    const uint64_t* ops_off;    // per pair word offset into ops
    uint32_t        n_pp;
    int32_t         R;
    int32_t         match, mismatch, gap, bias;
    int32_t         opt;
    const AlphaInfo* alpha;
};

struct DevLoader {
    const Chunk* base;
    __device__ __forceinline__ Chunk operator()(uint64_t idx) const {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + idx));
        return Chunk{v.x, v.y, v.z, v.w};
    }
    __device__ __forceinline__ void prefetch(uint64_t idx) const {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(base + idx));
    }
};

constexpr int TB_THREADS = 128;

template <int K, bool LOCAL>
__global__ void __launch_bounds__(TB_THREADS)
short16_traceback_kernel(const TbArgs A)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pp = t >> 1;
    const int half = (int)(t & 1u);
    if (pp >= A.n_pp || A.alpha->too_many) return;
    const PPDesc d = A.pps[pp];
    if (half && d.b == d.a) return;                       // singleton: the high half is a duplicate
    const uint32_t pair = half ? d.b : d.a;
    const Chunk* rec = A.codes + A.code_off[pp];
    PairView v{rec, LOCAL ? A.rowbest + (size_t)pp * A.R * 32u : nullptr,
               A.pat + A.pat_off[pair], A.txt + A.txt_off[pair],
               d.m, d.n, num_chunks(d.n, Geo<K>::CS), A.R, half, A.match, A.mismatch, A.gap, A.bias, A.opt, short16_rmagic(A.R)};
    OpsSink sink(A.ops ? A.ops + A.ops_off[pair] : nullptr);
    PairResult res;
    if (LOCAL) walk_local<Short16<K>>(v, DevLoader{rec}, sink, res);
    else walk_global<Short16<K>>(v, DevLoader{rec}, sink, res);
    sink.flush();
    res.path = 1;
    A.results[pair] = res;
}

} // namespace b2a
