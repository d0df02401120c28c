// wide32.cuh -- int32 kernel family: intra-pair tiled anti-diagonal wavefront for pairs the s16x2
// record cannot hold (long patterns, general byte alphabets, large scores).
//
// Replaces the same reference loops as short16 (hw2.cpp:138-156 / :205-231) for ONE pair spread over
// many warps: the DP matrix is cut into BANDS of 128 pattern rows (32 lanes x 4 rows); a warp sweeps
// its band along the text as a skewed wavefront (__shfl_up_sync carries the in-warp diagonal
// dependency) and hands the band's bottom row to the band below through L2 as SELF-VALIDATING 64-bit
// entries {H, tag}: value and validity travel in one atomic word, so neither side needs a fence or a
// separate progress flag (an earlier st.release/ld.acquire counter design spent 60 % of a single long
// pair's time in membar/long-scoreboard stalls, profiles/r01_ncu_wide32_c4.md).  The consumer loads
// its column of the next 32-column block one block ahead and only re-polls if the tag is not there yet.
// tag = epoch * 2^20 + band + 1: the two boundary buffers of a pair alternate between bands, the epoch
// changes with every launch, the host clears the buffers when the epoch wraps or the buffer moves.
// Bands are handed out by a ticket counter in (band, pair) order, so the band a warp waits for was
// always claimed earlier by a warp that is already running: no cooperative launch is needed and
// nothing can deadlock.
// The traceback record is the same delta/anchor chunk format as short16 (b2a_format.h, Wide32<K>),
// written with the same ring-arithmetic word trick; K = 32 stores raw deltas and covers any scoring.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "b2a_format.h"
#include "short16_fill.cuh"
#include "walk_warp.cuh"

namespace b2a {

struct WidePair {
    uint64_t pat_off, txt_off;      // byte offsets of the pair's sequences
    uint64_t code_off;              // first chunk of the pair's record
    uint64_t bound_off;             // entry index of the pair's 2 boundary rows (each bound_stride long); multi-band pairs only
    uint64_t rowbest_off;           // uint32 index, nbands*128 entries
    uint32_t m, n, pair, nbands;
    uint32_t bound_stride, pad;
};
struct WideTask { uint32_t wp, band; };

struct WideArgs {
    const uint8_t*  pat;
    const uint8_t*  txt;
    const WidePair* pairs;
    const WideTask* tasks;
    uint32_t        n_tasks;
    uint32_t*       ticket;
    Chunk*          codes;
    uint64_t*       bound;          // tagged boundary entries: (tag << 32) | (uint32) H
    uint32_t*       rowbest;
    int32_t*        final_score;    // per wide pair: H(m, n) (global mode)
    int32_t         match, mismatch, gap;
    uint32_t        radix;
    uint32_t        epoch_tag;      // epoch << 20
    unsigned long long* debug;      // optional (B2A_WIDE_DEBUG): per task {claimed, first block staged, first block done, band done} in ns
    const AlphaInfo* alpha;         // used by the ALPHA4 variant only
};

constexpr int WIDE_WARPS = 4;
#ifndef WIDE_MID_UNROLL
#define WIDE_MID_UNROLL 32     // unroll of the F-2 middle steps of a delta word (tuned on config 4 / config 5, scripts/bench_long.py)
#endif
constexpr int MID_UNROLL = WIDE_MID_UNROLL;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));       // no "memory" clobber: nothing else is ordered by it
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v));
}
// Waits until this lane's tagged boundary entry (if it `need`s one) carries `tag`; `v` is a copy loaded a block
// earlier.  WARP-UNIFORM on purpose: every lane leaves the loop in the same iteration (__all_sync).  A per-lane
// polling loop left the warp split into sub-warps that then executed the whole next block of shuffle-synchronised
// steps one group after the other (3.5x slower first block of every band, measured with B2A_WIDE_DEBUG stamps).
// Re-polls back off so that hundreds of waiting bands of one long pair do not flood L2.
__device__ __forceinline__ uint32_t bound_wait(const uint64_t* p, uint64_t v, uint32_t tag, bool need) {
    uint32_t ns = 32;
    for (;;) {
        const bool ok = !need || (uint32_t)(v >> 32) == tag;
        if (__all_sync(0xFFFFFFFFu, ok)) break;
        __nanosleep(ns); ns = min(ns * 2u, 256u);
        if (!ok) v = ld_relaxed_u64(p);
    }
    return (uint32_t)v;
}
__device__ __forceinline__ uint32_t prmt32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// K: delta width; LOCAL: Smith-Waterman; STORE: write the traceback record; ALPHA4: pattern alphabet
// has <= 4 symbols and scores fit int8, so the substitution score is one PRMT from a 4-entry table.
template <int K, bool LOCAL, bool STORE, bool ALPHA4>
__global__ void __launch_bounds__(WIDE_WARPS * 32)
wide32_fill_kernel(const WideArgs A)
{
    using FM = Wide32<K>;
    constexpr int R = WIDE_R, F = FM::F, CS = FM::CS, CPB = 32 / CS;   // chunks per 32-step block
    __shared__ uint2 s_ring[WIDE_WARPS][64];       // per 1-based column j (slot j & 63): {text entry, H(top row, j)}
    __shared__ int32_t s_out[WIDE_WARPS][32];      // bottom-row values lane 31 produced during the current block
    __shared__ uint32_t s_tbl4[256];
    uint8_t sym[4] = {0, 0, 0, 0};
    if (ALPHA4) {
        const int nsym = A.alpha->nsym;
#pragma unroll
        for (int c = 0; c < 4; ++c) sym[c] = A.alpha->sym[c];
        for (int b = threadIdx.x; b < 256; b += blockDim.x) {
            uint32_t w = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int sc = (c < nsym && sym[c] == (uint8_t)b) ? A.match : A.mismatch;
                w |= ((uint32_t)sc & 0xFFu) << (8 * c);
            }
            s_tbl4[b] = w;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint2* ring = s_ring[warp];
    int32_t* outb = s_out[warp];
    const int gap = A.gap;
    const uint32_t radix = A.radix, g32 = (uint32_t)gap;
    uint32_t geo = 0, bpow = 1;
    if (K < 32) {
#pragma unroll
        for (int t = 0; t < F; ++t) geo = geo * radix + 1u;
#pragma unroll
        for (int t = 0; t < F - 1; ++t) bpow *= radix;
    }
    const uint32_t negGc = 0u - g32 * geo, negBpow = 0u - bpow, radm1 = radix - 1u;

    for (;;) {
        uint32_t tk = 0;
        if (lane == 0) tk = atomicAdd(A.ticket, 1u);
        tk = __shfl_sync(0xFFFFFFFFu, tk, 0);
        if (tk >= A.n_tasks) break;
        const WideTask task = A.tasks[tk];
        const WidePair wp = A.pairs[task.wp];
        const uint32_t m = wp.m, n = wp.n, band = task.band;
        const uint8_t* pp = A.pat + wp.pat_off;
        const uint8_t* tt = A.txt + wp.txt_off;
        const uint32_t row0 = band * 32u * R + (uint32_t)lane * R;          // 0-based first row of this lane
        uint32_t pc[R];
        int32_t H[R], best[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i0 = row0 + r;
            if (ALPHA4) {
                uint32_t c = 0;
                if (i0 < m) { const uint8_t x = pp[i0];
#pragma unroll
                    for (int k = 1; k < 4; ++k) if (x == sym[k]) c = k; }
                pc[r] = c | ((8u | c) << 4) | ((8u | c) << 8) | ((8u | c) << 12);
            } else pc[r] = i0 < m ? (uint32_t)pp[i0] : 0xFFFFFFFFu;         // junk rows never match
            H[r] = LOCAL ? 0 : (int32_t)(i0 + 1) * gap;                      // hw2.cpp:125-130
            best[r] = 0;
        }
        const uint64_t* bin = A.bound + wp.bound_off + (uint64_t)((band + 1u) & 1u) * wp.bound_stride;
        uint64_t* bout = A.bound + wp.bound_off + (uint64_t)(band & 1u) * wp.bound_stride;
        const uint32_t tag_in = A.epoch_tag + band;                          // what the band above writes
        const uint64_t tag_out = (uint64_t)(A.epoch_tag + band + 1u) << 32;
        const bool has_next = band + 1 < wp.nbands;
        const uint32_t nblk = (n + 32u + 31u) / 32u;
        const uint32_t NC = num_chunks(n, CS);
        Chunk* rec = A.codes + wp.code_off;
        const int32_t top0 = LOCAL ? 0 : (int32_t)(band * 32u * R) * gap;    // H(top row, 0)
        int32_t dgn = top0;
        uint64_t nx = 0;                                                     // this lane's entry of the next block, loaded a block ahead
        if (band != 0 && lane >= 1 && (uint32_t)lane <= n) nx = ld_relaxed_u64(bin + lane);

        // text byte of 1-based column j, fetched one block ahead (0x100 = outside the text); the score-table
        // lookup waits until the byte is staged, so the global load has a whole block to arrive
        auto text_byte = [&](uint32_t j) -> uint32_t { return (j == 0 || j > n) ? 0x100u : (uint32_t)tt[j - 1]; };
        auto text_entry = [&](uint32_t x) -> uint32_t {                      // ring payload
            if (x & 0x100u) return ALPHA4 ? 0u : 0xFFFFFF00u;
            return ALPHA4 ? s_tbl4[x] : x;
        };
        uint32_t tnext = text_byte((uint32_t)lane);
        auto stamp = [&](int k) {
            if (A.debug && lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); A.debug[8ull * tk + k] = t; }
        };
        stamp(0);

        for (uint32_t kb = 0; kb < nblk; ++kb) {
            const uint32_t q0 = kb * 32u;
            if (kb == 1) stamp(4);
            if (kb == 2) stamp(6);
            // ---- stage columns q0 .. q0+31 of the text and of the band above into the ring ----
            int32_t bnd;
            const uint32_t jcol = q0 + (uint32_t)lane;
            if (band == 0) bnd = LOCAL ? 0 : (int32_t)jcol * gap;            // hw2.cpp:131-136
            else {
                const bool need = jcol >= 1 && jcol <= n;
                const int32_t got = (int32_t)bound_wait(bin + jcol, nx, tag_in, need);
                bnd = need ? got : top0;
                const uint32_t jn = jcol + 32u;
                if (jn <= n) nx = ld_relaxed_u64(bin + jn);
            }
            __syncwarp();
            if (kb == 0) stamp(1);
            if (kb == 1) stamp(5);
            ring[jcol & 63u] = make_uint2(text_entry(tnext), (uint32_t)bnd);
            __syncwarp();
            if (kb == 0) stamp(2);
            tnext = text_byte(q0 + 32u + (uint32_t)lane);                    // prefetch the next block's text
            // One wavefront step.  RAMP = false: every lane is inside its row range (steady state).  RAMP = true: lanes
            // outside 1 <= q - lane <= n keep their H frozen -- done with selects, NOT a branch: a branch per step stops
            // the scheduler from overlapping the shuffle / shared-memory latencies of neighbouring steps (2.4x slower).
            const bool steady = q0 >= 32u && q0 + 31u <= n;
            // HM: what the step contributes to the delta word's Horner sum: 0 = start it (S = H), 1 = S = S * 2^K + H, 2 = nothing
            auto step = [&](uint32_t q, uint32_t (&S)[R], auto hm_tag, auto ramp_tag, bool active) {
                constexpr bool RAMP = decltype(ramp_tag)::value;
                constexpr int HM = decltype(hm_tag)::value;
                int32_t up = __shfl_up_sync(0xFFFFFFFFu, H[R - 1], 1);
                const uint2 e = ring[(q - (uint32_t)lane) & 63u];   // column q-lane; lane 0 also needs it at q = 0 (H(top,0))
                if (lane == 0) up = (int32_t)e.y;
                int32_t dg = dgn, u = up;
                dgn = up;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int32_t s = ALPHA4 ? (int32_t)prmt32(e.x, 0u, pc[r]) : (pc[r] == e.x ? A.match : A.mismatch);
                    const int32_t ds = dg + s;
                    dg = H[r];
                    const int32_t a = __viaddmax_s32(H[r], gap, ds);
                    int32_t h = LOCAL ? __viaddmax_s32_relu(u, gap, a) : __viaddmax_s32(u, gap, a);
                    if (RAMP) h = active ? h : H[r];
                    if (LOCAL) best[r] = max(best[r], h);        // a frozen H is a border 0 or a value already counted
                    H[r] = h; u = h;
                }
                if (lane == 31) outb[q & 31u] = H[R - 1];        // bottom row of the band at column q - 31 (published per block)
                if (STORE && K < 32 && HM != 2) {
#pragma unroll
                    for (int r = 0; r < R; ++r) S[r] = HM == 0 ? (uint32_t)H[r] : S[r] * radix + (uint32_t)H[r];
                }
            };
            // The F steps of one delta word.  Deliberately NOT fully unrolled: fully unrolled the two flavours were 38 KB
            // of straight-line code, and every band of a long pair paid ~6 us of instruction-fetch stalls for its first
            // (cold) ramp block -- a cost that chains over all bands of the pair (B2A_WIDE_DEBUG timeline, scripts/band_exp.py).
            auto word_steps = [&](uint32_t qw, uint32_t (&S)[R], auto ramp_tag) {
                constexpr bool RAMP = decltype(ramp_tag)::value;
                auto act = [&](uint32_t q) { return !RAMP || (uint32_t)(q - lane - 1u) < n; };
                if (F == 1) { step(qw, S, std::integral_constant<int, 2>{}, ramp_tag, act(qw)); return; }
                step(qw, S, std::integral_constant<int, 0>{}, ramp_tag, act(qw));
#pragma unroll MID_UNROLL
                for (int f = 1; f < F - 1; ++f) step(qw + (uint32_t)f, S, std::integral_constant<int, 1>{}, ramp_tag, act(qw + (uint32_t)f));
                step(qw + (uint32_t)(F - 1), S, std::integral_constant<int, 2>{}, ramp_tag, act(qw + (uint32_t)(F - 1)));
            };

#pragma unroll 1
            for (int cb = 0; cb < CPB; ++cb) {
                uint32_t w0[R];
#pragma unroll
                for (int wi = 0; wi < 2; ++wi) {
                    uint32_t S[R], pre[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) pre[r] = K == 32 ? (0u - (uint32_t)H[r] - g32) : (uint32_t)H[r] * negBpow + negGc;
                    const uint32_t qw = q0 + (uint32_t)(cb * CS + wi * F);
                    // Both flavours are straight-line code of the same speed: the last two blocks of band b wait for the last
                    // block of band b-1, so a slow ramp flavour is paid once PER BAND on the critical path of a long pair.
                    if (steady) word_steps(qw, S, std::false_type{});
                    else word_steps(qw, S, std::true_type{});
                    if (kb == 0 && cb == CPB - 1 && wi == 1) stamp(3);
                    if (STORE) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint32_t w = (K < 32 && F > 1 ? S[r] * radm1 : 0u) + (uint32_t)H[r] + pre[r];
                            if (wi == 0) w0[r] = w;
                            else {
                                const uint32_t c = kb * CPB + cb;
                                if (c < NC) {
                                    const uint4 v = make_uint4(w0[r], w, 0u, (uint32_t)H[r]);
                                    *reinterpret_cast<uint4*>(&rec[(((uint64_t)band * R + r) * NC + c) * 32u + lane]) = v;
                                }
                            }
                        }
                    }
                }
            }
            // publish the 32 bottom-row values of this block as tagged entries: ONE coalesced 64-bit store per lane
            // per block instead of a store per step (the per-step stores cost 8 % on the 120 x 100 kb batch)
            if (has_next) {
                __syncwarp();
                const uint32_t jc = q0 + (uint32_t)lane - 31u;                // lane 31 was at this column at step q0 + lane
                if (jc - 1u < n) st_relaxed_u64(bout + jc, tag_out | (uint32_t)outb[lane]);
                __syncwarp();
            }
        }
        stamp(7);
        if (LOCAL) {
#pragma unroll
            for (int r = 0; r < R; ++r) A.rowbest[wp.rowbest_off + ((uint64_t)band * R + r) * 32u + lane] = (uint32_t)best[r];
        } else if (band + 1 == wp.nbands) {
            const uint32_t ib = (m - 1u) - band * 32u * R;
            if ((uint32_t)lane == ib / R) {
                const uint32_t rm = ib % R;
                int32_t v = H[0];
#pragma unroll
                for (int r = 1; r < R; ++r) if (rm == (uint32_t)r) v = H[r];
                A.final_score[task.wp] = v;
            }
        }
        __syncwarp();
    }
}

// ---- traceback over the wide32 record: one thread per pair ----
struct WideTbArgs {
    const uint8_t*  pat;
    const uint8_t*  txt;
    const WidePair* pairs;
    uint32_t        n_wide;
    const Chunk*    codes;
    const uint32_t* rowbest;
    const int32_t*  final_score;
    PairResult*     results;
    uint32_t*       ops;
    const uint64_t* ops_off;
    int32_t         match, mismatch, gap;
    int32_t         score_only;
    int32_t         opt;            // bit 2 (value 4): hw4 tie order d > u > l, overlap := hw4's distance
};

struct WideLoader {
    const Chunk* base;
    __device__ __forceinline__ Chunk operator()(uint64_t idx) const {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + idx));
        return Chunk{v.x, v.y, v.z, v.w};
    }
    __device__ __forceinline__ void prefetch(uint64_t idx) const {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(base + idx));
    }
};

constexpr int WIDE_TB_WARPS = 2;

// One WARP per pair (walk_warp.cuh); score-only batches just gather the scores the fill kernel left.
template <int K, bool LOCAL>
__global__ void __launch_bounds__(WIDE_TB_WARPS * 32)
wide32_traceback_kernel(const WideTbArgs A)
{
    using FM = Wide32<K>;
    const int lane = (int)(threadIdx.x & 31u);
    const uint32_t t = blockIdx.x * WIDE_TB_WARPS + (threadIdx.x >> 5);
    if (t >= A.n_wide) return;
    const WidePair wp = A.pairs[t];
    PairResult res;
    if (A.score_only) {
        res = PairResult{0, 0, 0, 0, 0, 0, 0, 2};
        if (LOCAL) {
            int M = 0;
            for (uint32_t k = (uint32_t)lane; k < wp.nbands * 32u * WIDE_R; k += 32u) {   // k = (band*R + r)*32 + L; rows beyond m are junk
                const uint32_t L = k & 31u, br = k >> 5, r = br % WIDE_R, band = br / WIDE_R;
                if (band * 32u * WIDE_R + L * WIDE_R + r >= wp.m) continue;
                M = max(M, (int)A.rowbest[wp.rowbest_off + k]);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) M = max(M, __shfl_xor_sync(0xFFFFFFFFu, M, o));
            res.score = M;
        } else { res.score = (wp.m && wp.n) ? A.final_score[t] : (int32_t)(wp.m + wp.n) * A.gap; res.end_i = wp.m; res.end_j = wp.n; }
        if (lane == 0) A.results[wp.pair] = res;
        return;
    }
    const Chunk* rec = A.codes + wp.code_off;
    const PairView v{rec, A.rowbest + wp.rowbest_off, A.pat + wp.pat_off, A.txt + wp.txt_off,
                     wp.m, wp.n, num_chunks(wp.n, FM::CS), WIDE_R, 0, A.match, A.mismatch, A.gap, 0, A.opt, 0};
    const WideLoader ld{rec};
    WarpOpsSink sink(A.ops ? A.ops + A.ops_off[wp.pair] : nullptr, lane == 0);
    uint32_t i, j, nops = 0, mism = 0;
    int best = 0;
    if (LOCAL) {
        int M; uint32_t bi, bj;
        warp_find_local_end<FM>(v, ld, M, bi, bj);
        res.score = M; res.end_i = bi; res.end_j = bj;
        i = bi; j = bj;
        if (M != 0) warp_walk<FM, WideLoader, true>(v, ld, sink, i, j, nops, best, mism);
    } else {
        i = wp.m; j = wp.n;
        res.end_i = i; res.end_j = j;
        res.score = (i && j) ? A.final_score[t] : (int32_t)(i + j) * A.gap;               // hw2.cpp:186 / borders :125-136
        warp_walk<FM, WideLoader, false>(v, ld, sink, i, j, nops, best, mism);
        sink.put_run(OP_D, i); nops += i; i = 0;                                           // column 0 holds 'u' (hw2.cpp:128)
        sink.put_run(OP_I, j); nops += j; j = 0;                                           // row 0 holds 'l'    (hw2.cpp:134)
    }
    sink.flush();
    res.start_i = i; res.start_j = j; res.n_ops = nops; res.path = 2;
    res.overlap = (!LOCAL && (A.opt & 4)) ? (int)(nops - (wp.m + wp.n - nops) + mism) : best;       // hw4.cpp:141-152
    if (lane == 0) A.results[wp.pair] = res;
}

} // namespace b2a
