// traceback.cuh -- one thread per pair walks the short16 record.
//
// Replaces hw2.cpp:158-188 (NW) / hw2.cpp:235-263 (SW) and overlapLongestExactMatch (hw2.cpp:267-278)
// for every pair of the batch.  The direction of a cell is decided by the reference's own comparisons
// (hw2.cpp:145-153 NW, hw2.cpp:214-222 SW) on the exact H of the cell and its three neighbours.
//
// The walker keeps the two DP rows it stands between RESIDENT IN REGISTERS: per row the 48-bit string of
// the current chunk's deltas, the bit offset of the current column and the exact H there.  A move to the
// left costs two field extractions and no memory access; a chunk is loaded only when a row's column
// crosses a chunk edge (every CS columns) or when the path climbs a row (the row above becomes the
// current row as it is, one new chunk is sought for the row above that).  A 150 x 1000 global path of
// ~1150 steps touches ~230 chunks that way; the first version of this kernel re-sought both rows at every
// step (~2000 16-byte loads per pair, 7.6 KB of DRAM sectors, bound by L2 request throughput at 5.9 ms
// per million pairs).  Text bytes come through an aligned 4-byte window (one load per 4 columns).
// The generic cursor walkers in b2a_format.h stay the executable specification (CPU host model, wide32 warp walker).
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
    const int4*     endcell;    // local mode, per pair: {score, end_i, end_j, 0} from the fill kernel
    PairResult*     results;    // indexed by pair
    uint32_t*       ops;        // packed 2-bit ops, may be null
    const uint64_t* ops_off;    // per pair word offset into ops
    uint32_t        n_pp;
    int32_t         R;
    int32_t         match, mismatch, gap, bias;
    int32_t         opt;
    int32_t         tie_hw4;    // global mode: tie order d > u > l (hw4.cpp:37-46) and overlap := hw4's distance
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

// one pair's view of its half of a pair-pair record
template <int K>
struct S16View {
    using FM = Short16<K>;
    const Chunk* rec;
    uint32_t NC, R, rmagic, stride;    // stride = 32R: distance between consecutive chunks of a row
    uint32_t sel_lo, sel_hi;     // PRMT selectors of this pair's 16-bit half
    int gap;
    uint32_t pf;                 // L2 prefetch bits: 1 / 4 = rows above when the path climbs (near / far), 2 = the chunk to the left of every loaded chunk

    // first chunk of DP row i >= 1 (its chunks are 32R apart) and the lane that owned the row in the fill
    // (step q = j + L; no runtime division, see row_slot)
    __device__ __forceinline__ const Chunk* row_base(uint32_t i, uint32_t& L) const {
        const uint32_t x = i - 1u;
        L = (x * rmagic) >> 16;
        return rec + x;
    }
    // field string of a chunk (step rem at bit K*(CS-1-rem)) and its anchor
    __device__ __forceinline__ uint64_t unpack(const uint4& ch, int& anchor) const {
        const uint32_t lo = __byte_perm(ch.z, ch.y, sel_lo);              // (w1.half << 16) | w2.half
        const uint32_t hi = __byte_perm(ch.x, 0u, sel_hi);                // w0.half
        anchor = (int)__byte_perm(ch.w, 0u, sel_hi);
        return ((uint64_t)hi << 32) | lo;
    }
    // (tried and lost with the first walker, per 1 M pairs: prefetch.global.L1 of the rows above 6.2 -> 7.9 ms NW / 2.7 -> 6.2 ms SW;
    //  ld.cg instead of ld.nc for the chunks 6.2 -> 7.2 ms NW; ld.cs 6.2 -> 8.9 ms)
};

#ifndef TB_MIN_CTAS
#define TB_MIN_CTAS 1
#endif

// one DP row of the record as the walker holds it
template <int K>
struct TbRow {
    using FM = Short16<K>;
    static constexpr int BITS = K * FM::CS;      // 48
    uint64_t X;          // field string of the current chunk: step rem at bit K*(CS-1-rem)
    const Chunk* p;      // the current chunk; nullptr on the virtual border row 0
    uint32_t c;          // its chunk column
    int off, H;          // bit offset of the current column's field; exact (biased) H at the current column

    __device__ __forceinline__ int D() const { return (int)((uint32_t)(X >> off) & FM::MASK); }
    // the chunk to the left of the current one will be needed CS columns from now: ask L2 for it (the record lives in DRAM)
    __device__ __forceinline__ void prefetch_left(const S16View<K>& v) const {
        if (c > 0u) asm volatile("prefetch.global.L2 [%0];" :: "l"(p - v.stride));
    }
    // row i >= 1 at column j
    __device__ __forceinline__ void seek(const S16View<K>& v, uint32_t i, uint32_t j) {
        uint32_t L;
        const Chunk* rb = v.row_base(i, L);
        const uint32_t q = j + L;
        c = q / (uint32_t)FM::CS;
        const int rem = (int)(q - c * (uint32_t)FM::CS);
        p = rb + (size_t)c * v.stride;
        int anchor;
        X = v.unpack(__ldg(reinterpret_cast<const uint4*>(p)), anchor);
        off = K * (FM::CS - 1 - rem);
        H = anchor - (FM::CS - 1 - rem) * v.gap - field_sum64<K>(X & ((1ull << off) - 1ull));
        if (v.pf & 2u) prefetch_left(v);
    }
    // the border row 0 at column j: H = bias + j*gap with D = 0 (global, hw2.cpp:131-136) / H = 0 with D = -gap (local, hw2.cpp:196-197)
    __device__ __forceinline__ void border(bool local, int bias, int gap, uint32_t j) {
        p = nullptr; c = 0; off = 0;
        X = 0;
        if (local) {
#pragma unroll
            for (int f = 0; f < FM::CS; ++f) X |= (uint64_t)(uint32_t)(-gap) << (K * f);
        }
        H = local ? 0 : bias + (int)j * gap;
    }
    // one column to the left (the caller guarantees the column exists)
    __device__ __forceinline__ void left(const S16View<K>& v) {
        H -= D() + v.gap;
        off += K;
        if (off == BITS) {
            off = 0;
            if (p != nullptr) {
                if (c > 0u) { --c; p -= v.stride; int anchor; X = v.unpack(__ldg(reinterpret_cast<const uint4*>(p)), anchor); if (v.pf & 2u) prefetch_left(v); }
                else off = BITS - K;                     // step 0 of the row: nothing further left is ever read
            }
        }
    }
};

// text bytes through an aligned 4-byte window (the device buffers are 256-byte aligned and padded, so the aligned word exists)
struct TextWindow {
    const uint8_t* t;
    uint32_t w;
    uintptr_t at;        // address of the word in w; 1 = none yet
    __device__ __forceinline__ explicit TextWindow(const uint8_t* t_) : t(t_), w(0), at(1) {}
    __device__ __forceinline__ uint32_t get(uint32_t idx) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(t + idx), base = a & ~(uintptr_t)3;
        if (base != at) { at = base; w = __ldg(reinterpret_cast<const uint32_t*>(base)); }
        return (w >> (8u * (uint32_t)(a & 3u))) & 0xFFu;
    }
};

template <int K, bool LOCAL>
__global__ void __launch_bounds__(TB_THREADS, TB_MIN_CTAS)
short16_traceback_kernel(const TbArgs A)
{
    using FM = Short16<K>;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pp = t >> 1;
    const int half = (int)(t & 1u);
    if (pp >= A.n_pp || A.alpha->too_many) return;
    const PPDesc d = A.pps[pp];
    if (half && d.b == d.a) return;                       // singleton: the high half is a duplicate
    const uint32_t pair = half ? d.b : d.a;
    const uint8_t* __restrict__ P = A.pat + A.pat_off[pair];
    TextWindow T(A.txt + A.txt_off[pair]);
    const int gap = A.gap, match = A.match, mismatch = A.mismatch;
    const uint32_t mh = pp_dim(d.m, half), nh = pp_dim(d.n, half);   // this pair's own shape; the record is laid out for max(n)
    const S16View<K> v{A.codes + A.code_off[pp], num_chunks(pp_max(d.n), FM::CS), (uint32_t)A.R, (uint32_t)short16_rmagic(A.R), 32u * (uint32_t)A.R,
                       half ? 0x7632u : 0x5410u, half ? 0x4432u : 0x4410u, gap, (uint32_t)A.opt};
    uint32_t* out = A.ops ? A.ops + A.ops_off[pair] : nullptr;

    uint32_t i, j, nops = 0, word = 0, fill = 0, wpos = 0;
    int cur = 0, best = 0;
    const bool hw4 = !LOCAL && A.tie_hw4 != 0;
    uint32_t mism = 0;                                    // hw4: 'M' columns whose bases differ
    PairResult res;
    res.path = 1;
    auto put = [&](uint32_t op) {
        word |= op << (2u * fill);
        if (++fill == 16u) { if (out) out[wpos] = word; ++wpos; word = 0; fill = 0; }
    };
    auto put_run = [&](uint32_t op, uint32_t cnt) {
        while (cnt) {
            const uint32_t take = min(cnt, 16u - fill);
            if (op) word |= ((op * 0x55555555u) & (take == 16u ? 0xFFFFFFFFu : ((1u << (2u * take)) - 1u))) << (2u * fill);
            fill += take; cnt -= take;
            if (fill == 16u) { if (out) out[wpos] = word; ++wpos; word = 0; fill = 0; }
        }
    };

    if (LOCAL) {
        const int4 e = A.endcell[pair];
        res.score = e.x; res.end_i = (uint32_t)e.y; res.end_j = (uint32_t)e.z;
        i = res.end_i; j = res.end_j;
        if (e.x == 0) { i = 0; j = 0; }                   // hw2.cpp:202-203: best cell stays (0,0), empty alignment
    } else {
        i = mh; j = nh;
        res.end_i = i; res.end_j = j;
    }
    TbRow<K> ra, rb;                                      // rows i and i-1 at column j
    ra.border(LOCAL, A.bias, gap, j); rb = ra;
    if (i > 0u && j > 0u) {
        ra.seek(v, i, j);
        if (i > 1u) rb.seek(v, i - 1u, j); else rb.border(LOCAL, A.bias, gap, j);
    }
    if (!LOCAL) res.score = ra.H - A.bias;                // hw2.cpp:186
    uint32_t pc = i > 0u ? P[i - 1u] : 0u;

    while (i > 0u && j > 0u) {
        const int H = ra.H;
        if (LOCAL && H == 0) break;                                                          // hw2.cpp:239
        const int Hl = H - ra.D() - gap, Hu = rb.H, Hd = Hu - rb.D() - gap;                  // H(i,j-1), H(i-1,j), H(i-1,j-1)
        const bool eq = pc == T.get(j - 1u);
        const int dv = Hd + (eq ? match : mismatch);                                         // hw2.cpp:142 / :208
        uint32_t op;
        if (LOCAL) op = H == dv ? OP_M : (H == Hu + gap ? OP_D : OP_I);                      // hw2.cpp:214-222 (H != 0 here)
        else if (!hw4) { op = OP_M; int val = dv; if (Hl + gap > val) { val = Hl + gap; op = OP_I; } if (Hu + gap > val) op = OP_D; }   // hw2.cpp:145-153
        else { op = OP_M; int val = dv; if (Hu + gap > val) { val = Hu + gap; op = OP_D; } if (Hl + gap > val) op = OP_I; }            // hw4.cpp:37-46
        mism += (op == OP_M && !eq);
        cur = (op == OP_M && eq && pc != (uint32_t)'-') ? cur + 1 : 0;                       // hw2.cpp:267-278
        best = max(best, cur);
        put(op); ++nops;
        if (op != OP_D) {                                 // 'M' and 'I' move one column left
            --j;
            rb.left(v);
            if (op == OP_I) ra.left(v);
        }
        if (op != OP_I) {                                 // 'M' and 'D' climb one row: the row above becomes the current row as it stands
            ra = rb; --i;
            if (i > 0u) {
                pc = P[i - 1u];
                if (j > 0u) { if (i > 1u) rb.seek(v, i - 1u, j); else rb.border(LOCAL, A.bias, gap, j); }
                // The chunk the NEXT climbs will seek is predictable (same column, one and two rows further up): its DRAM latency then
                // overlaps the steps in between instead of stalling the warp at the climb.
                if ((v.pf & 5u) && j > 0u) {
                    auto pf_row = [&](uint32_t k, uint32_t back) {                 // row i - k at column j - back
                        if (i > k) {
                            uint32_t L;
                            const Chunk* rbase = v.row_base(i - k, L);
                            const uint32_t col = j > back ? j - back : 1u;
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(rbase + (size_t)((col + L) / (uint32_t)FM::CS) * v.stride));
                        }
                    };
                    if (v.pf & 1u) { pf_row(2u, 0u); pf_row(3u, 0u); }           // staircase paths (a climb every few columns)
                    if (v.pf & 4u) { pf_row(4u, 2u); pf_row(6u, 4u); }           // diagonal paths (a climb every step)
                }
            }
        }
    }
    if (!LOCAL) {
        put_run(OP_D, i); nops += i; i = 0;                                                  // column 0 holds 'u' (hw2.cpp:128)
        put_run(OP_I, j); nops += j; j = 0;                                                  // row 0 holds 'l'    (hw2.cpp:134)
    }
    if (fill && out) out[wpos] = word;
    res.start_i = i; res.start_j = j; res.n_ops = nops;
    res.overlap = hw4 ? (int)(nops - (mh + nh - nops) + mism) : best;      // gap columns = n_ops - M columns, M columns = m + n - n_ops
    A.results[pair] = res;
}

} // namespace b2a
