// affine32.cuh -- score-only 3-state affine global alignment (sm_100a): the distance stage of the
// reference's centre-star MSA, Multiple_Sequence_Alignment/hw3.cpp:23-98 called from hw3.cpp:231-241.
//
// The reference recurrences (note: no E<->F transitions, V = "ends in a (mis)match"):
//     V[i][j] = max(V, F, E)[i-1][j-1] + s(i,j)                         hw3.cpp:57-68
//     F[i][j] = max(V[i-1][j] + Go + Ge, F[i-1][j] + Ge)                hw3.cpp:70-75
//     E[i][j] = max(V[i][j-1] + Go + Ge, E[i][j-1] + Ge)                hw3.cpp:77-82
//     borders V[0][0] = 0, F[i][0] = Go + Ge(i-1), E[0][j] = Go + Ge(j-1), all else NEG = INT_MIN/2   hw3.cpp:40-53
//     score   = max(V, F, E)[m][n]                                      hw3.cpp:86-98
// are evaluated in the same int32 arithmetic with the same sentinel, so every cell (also the
// sentinel-tainted ones) holds exactly the reference's value; only the score is wanted here, so no
// tie-breaking is involved.
//
// Same execution scheme as wide32.cuh: bands of 128 rows (32 lanes x 4 rows), a warp sweeps its band
// as a skewed wavefront, bands of one pair are chained through boundary rows of self-validating
// tagged 64-bit entries in L2 (three per column: Vg, F, M3), bands are claimed from a ticket counter
// in (band, pair) order inside one launch.  Per cell the lane keeps  Vg = V + Go + Ge,  E  and  M3 = max(V, F, E); F only travels
// downwards inside the step.  6 integer instructions per cell:
//     PRMT|SEL (s)   IADD (V = M3diag + s)   VIADDMNMX (F)   VIADDMNMX (E)   IADD (Vg)   VIMNMX3 (M3)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "b2a_format.h"
#include "wide32.cuh"

namespace b2a {

struct AffineArgs {
    const uint8_t*  pat;
    const uint8_t*  txt;
    const WidePair* pairs;          // bound_off indexes 2 buffers x 3 planes x bound_stride tagged entries
    const WideTask* tasks;
    uint32_t        n_tasks;
    uint32_t*       ticket;
    uint64_t*       bound;
    int32_t*        final_score;    // per pair: max(V, F, E)[m][n]
    int32_t         match, mismatch, gopen, gext;
    uint32_t        epoch_tag;
    const AlphaInfo* alpha;         // ALPHA4 variant only
    uint4*          codes;          // TRACE variant: 4-bit trace codes, 32 cells per 16 bytes, at WidePair::code_off
};

// Trace code of a cell (TRACE variant).  Every decision hw3's traceback reads is a property of ONE cell's (V, F, E):
//   bits 0-1  which of V, F, E is the largest, the later one winning only when strictly larger (hw3.cpp:59-68 as seen from
//             the cell diagonally below, and the final-state pick hw3.cpp:86-98): what traceV of the next cell will say
//   bit 2     F + Ge > V + Go + Ge: the vertical gap of the cell below extends (traceF = 1, hw3.cpp:70-75)
//   bit 3     E + Ge > V + Go + Ge: the horizontal gap of the cell to the right extends (traceE = 1, hw3.cpp:77-82)
// Layout: chunk ((band*4 + r)*nblk + kb)*32 + L holds the 32 steps of block kb of row r of lane L, step f in nibble f.
__device__ __forceinline__ uint32_t affine_code(int32_t v, int32_t f, int32_t e, int32_t vg, int32_t ge) {
    uint32_t b3 = f > v ? 1u : 0u;
    if (e > max(v, f)) b3 = 2u;
    return b3 | (f + ge > vg ? 4u : 0u) | (e + ge > vg ? 8u : 0u);
}

#ifndef AFFINE_UNROLL
#define AFFINE_UNROLL 32
#endif
constexpr int AFF_UNROLL = AFFINE_UNROLL;
constexpr int32_t AFFINE_NEG = INT32_MIN / 2;        // hw3.cpp:16

// R rows per lane: 4 with TRACE (the codes of 4 rows x 32 steps already take 16 registers), 8 for score-only runs, where the
// per-step overhead (3 shuffles, ring read, boundary staging) is then shared by twice the cells (+20 % on the 120-pair batch).
constexpr int AFFINE_R_SCORE = 8;

template <bool ALPHA4, bool TRACE, int R>
__global__ void __launch_bounds__(WIDE_WARPS * 32)
affine32_score_kernel(const AffineArgs A)
{
    __shared__ uint4 s_ring[WIDE_WARPS][64];       // per 1-based column j (slot j & 63): {text entry, Vg, F, M3 of the row above the band}
    __shared__ uint4 s_out[WIDE_WARPS][32];        // {Vg, F, M3} of the band's bottom row, produced by lane 31 during the current block
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
    uint4* ring = s_ring[warp];
    uint4* outb = s_out[warp];
    const int32_t ge = A.gext, go = A.gopen, goe = A.gopen + A.gext, NEG = AFFINE_NEG;

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
        const uint32_t row0 = band * 32u * R + (uint32_t)lane * R;          // 0-based first row of this lane = 1-based row above it
        uint32_t pc[R];
        int32_t Vg[R], E[R], M3[R];                                          // state at the lane's current column
        int32_t Fk;                                                          // F of the lane's last row
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t i0 = row0 + r;                                    // 1-based row i = i0 + 1
            if (ALPHA4) {
                uint32_t c = 0;
                if (i0 < m) { const uint8_t x = pp[i0];
#pragma unroll
                    for (int k = 1; k < 4; ++k) if (x == sym[k]) c = k; }
                pc[r] = c | ((8u | c) << 4) | ((8u | c) << 8) | ((8u | c) << 12);
            } else pc[r] = i0 < m ? (uint32_t)pp[i0] : 0xFFFFFFFFu;         // junk rows never match
            Vg[r] = NEG + goe;                                               // V[i][0] = NEG            hw3.cpp:43
            E[r]  = NEG;                                                     // E[i][0] = NEG            hw3.cpp:46
            M3[r] = go + ge * (int32_t)i0;                                   // F[i][0] = Go + Ge(i-1)   hw3.cpp:44
        }
        Fk = M3[R - 1];
        const uint64_t plane = wp.bound_stride;
        const uint64_t* bin = A.bound + wp.bound_off + (uint64_t)((band + 1u) & 1u) * 3u * plane;
        uint64_t* bout = A.bound + wp.bound_off + (uint64_t)(band & 1u) * 3u * plane;
        const uint32_t tag_in = A.epoch_tag + band;
        const uint64_t tag_out = (uint64_t)(A.epoch_tag + band + 1u) << 32;
        const bool has_next = band + 1 < wp.nbands;
        const uint32_t nblk = (n + 32u + 31u) / 32u;
        // max(V, F, E) of the row above the band at column 0: V[0][0] = 0 for the first band, else F[i][0]
        const int32_t top0 = band == 0 ? 0 : go + ge * (int32_t)(band * 32u * R - 1u);
        int32_t dgn = row0 == 0 ? 0 : go + ge * (int32_t)(row0 - 1u);       // M3[row above][0]; lane 0 reloads it from the ring at q = 0
        uint64_t nx0 = 0, nx1 = 0, nx2 = 0;                                   // this lane's entries of the next block, loaded a block ahead
        if (band != 0 && lane >= 1 && (uint32_t)lane <= n) {
            nx0 = ld_relaxed_u64(bin + lane); nx1 = ld_relaxed_u64(bin + plane + lane); nx2 = ld_relaxed_u64(bin + 2u * plane + lane);
        }

        auto text_byte = [&](uint32_t j) -> uint32_t { return (j == 0 || j > n) ? 0x100u : (uint32_t)tt[j - 1]; };   // one block ahead
        auto text_entry = [&](uint32_t x) -> uint32_t {
            if (x & 0x100u) return ALPHA4 ? 0u : 0xFFFFFF00u;
            return ALPHA4 ? s_tbl4[x] : x;
        };
        uint32_t tnext = text_byte((uint32_t)lane);

        for (uint32_t kb = 0; kb < nblk; ++kb) {
            const uint32_t q0 = kb * 32u;
            const uint32_t jcol = q0 + (uint32_t)lane;
            int32_t bVg, bF, bM;
            if (band == 0) {                                                 // row 0: V = F = NEG, E[0][j] = Go + Ge(j-1)   hw3.cpp:48-53
                bVg = NEG + goe; bF = NEG;
                bM = jcol == 0 ? 0 : go + ge * (int32_t)(jcol - 1u);
            } else {
                const bool need = jcol >= 1 && jcol <= n;
                bVg = (int32_t)bound_wait(bin + jcol, nx0, tag_in, need);
                bF  = (int32_t)bound_wait(bin + plane + jcol, nx1, tag_in, need);
                bM  = (int32_t)bound_wait(bin + 2u * plane + jcol, nx2, tag_in, need);
                if (!need) { bVg = NEG + goe; bF = NEG; bM = top0; }
                const uint32_t jn = jcol + 32u;
                if (jn <= n) { nx0 = ld_relaxed_u64(bin + jn); nx1 = ld_relaxed_u64(bin + plane + jn); nx2 = ld_relaxed_u64(bin + 2u * plane + jn); }
            }
            __syncwarp();
            ring[jcol & 63u] = make_uint4(text_entry(tnext), (uint32_t)bVg, (uint32_t)bF, (uint32_t)bM);
            __syncwarp();
            tnext = text_byte(q0 + 32u + (uint32_t)lane);

            const bool steady = q0 >= 32u && q0 + 31u <= n;
            uint32_t cw[R][4];                                               // TRACE: the block's codes of this lane's rows
            if (TRACE) {
#pragma unroll
                for (int r = 0; r < R; ++r) { cw[r][0] = cw[r][1] = cw[r][2] = cw[r][3] = 0u; }
            }
            // RAMP flavour freezes the lanes outside 1 <= q - lane <= n with selects, not a branch (see wide32.cuh)
            auto step = [&](uint32_t q, int fi, auto ramp_tag, bool active) {
                constexpr bool RAMP = decltype(ramp_tag)::value;
                int32_t uVg = __shfl_up_sync(0xFFFFFFFFu, Vg[R - 1], 1);
                int32_t uF  = __shfl_up_sync(0xFFFFFFFFu, Fk, 1);
                int32_t uM  = __shfl_up_sync(0xFFFFFFFFu, M3[R - 1], 1);
                const uint4 e = ring[(q - (uint32_t)lane) & 63u];
                if (lane == 0) { uVg = (int32_t)e.y; uF = (int32_t)e.z; uM = (int32_t)e.w; }
                int32_t dg = dgn;
                dgn = uM;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int32_t s = ALPHA4 ? (int32_t)prmt32(e.x, 0u, pc[r]) : (pc[r] == e.x ? A.match : A.mismatch);
                    const int32_t v = dg + s;                                // hw3.cpp:59-68
                    dg = M3[r];
                    const int32_t f  = __viaddmax_s32(uF, ge, uVg);          // hw3.cpp:70-75
                    const int32_t ee = __viaddmax_s32(E[r], ge, Vg[r]);      // hw3.cpp:77-82
                    const int32_t vg = v + goe;
                    const int32_t m3 = __vimax3_s32(v, f, ee);
                    if (TRACE) cw[r][fi >> 3] |= affine_code(v, f, ee, vg, ge) << (4 * (fi & 7));   // cells outside the matrix: never read
                    if (RAMP) { M3[r] = active ? m3 : M3[r]; Vg[r] = active ? vg : Vg[r]; E[r] = active ? ee : E[r]; }
                    else { M3[r] = m3; Vg[r] = vg; E[r] = ee; }
                    uVg = vg; uF = f;
                }
                if (RAMP) Fk = active ? uF : Fk; else Fk = uF;
                if (lane == 31) outb[q & 31u] = make_uint4((uint32_t)Vg[R - 1], (uint32_t)Fk, (uint32_t)M3[R - 1], 0u);
            };
            // small unroll on purpose: compact code keeps a band's first (cold) blocks cheap, see wide32.cuh
            if (steady) {
#pragma unroll AFF_UNROLL
                for (int f = 0; f < 32; ++f) step(q0 + (uint32_t)f, f, std::false_type{}, true);
            } else {
#pragma unroll AFF_UNROLL
                for (int f = 0; f < 32; ++f) step(q0 + (uint32_t)f, f, std::true_type{}, (uint32_t)(q0 + (uint32_t)f - lane - 1u) < n);
            }
            if (TRACE) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    A.codes[wp.code_off + (((uint64_t)band * R + r) * nblk + kb) * 32u + lane] = make_uint4(cw[r][0], cw[r][1], cw[r][2], cw[r][3]);
            }
            if (has_next) {                                                  // publish the block's bottom row, one coalesced store per plane
                __syncwarp();
                const uint32_t jc = q0 + (uint32_t)lane - 31u;
                if (jc - 1u < n) {
                    const uint4 o = outb[lane];
                    st_relaxed_u64(bout + jc, tag_out | o.x);
                    st_relaxed_u64(bout + plane + jc, tag_out | o.y);
                    st_relaxed_u64(bout + 2u * plane + jc, tag_out | o.z);
                }
                __syncwarp();
            }
        }
        if (band + 1 == wp.nbands) {                                         // hw3.cpp:86-98: max(V, F, E)[m][n]
            const uint32_t ib = (m - 1u) - band * 32u * R;
            if ((uint32_t)lane == ib / R) {
                const uint32_t rm = ib % R;
                int32_t v = M3[0];
#pragma unroll
                for (int r = 1; r < R; ++r) if (rm == (uint32_t)r) v = M3[r];
                A.final_score[task.wp] = v;
            }
        }
        __syncwarp();
    }
}

// ---- traceback over the trace codes: one WARP per pair, hw3.cpp:100-135 ----
// State machine of the reference (state V: emit a column of two bases, go diagonally, next state = traceV; state F: base over '-',
// go up, next state F iff traceF says "extended"; state E: '-' over base, go left).  Each state continues in one direction while one
// bit of the cells it passes stays set, so the 32 lanes read the codes of the next 32 cells of that direction and the leading run is
// accepted at once (same idea as walk_warp.cuh).  Ops use hw2's letters: M diagonal, D string1 base over '-', I '-' over string2 base.
struct AffineTbArgs {
    const WidePair* pairs;
    uint32_t        n_wide;
    const uint4*    codes;
    uint32_t*       n_ops;          // per pair (indexed by WidePair::pair)
    uint32_t*       ops;
    const uint64_t* ops_off;        // per pair word offset
};

__global__ void __launch_bounds__(WIDE_TB_WARPS * 32)
affine32_traceback_kernel(const AffineTbArgs A)
{
    constexpr int R = WIDE_R;
    const int lane = (int)(threadIdx.x & 31u);
    const uint32_t t = blockIdx.x * WIDE_TB_WARPS + (threadIdx.x >> 5);
    if (t >= A.n_wide) return;
    const WidePair wp = A.pairs[t];
    const uint32_t nblk = (wp.n + 32u + 31u) / 32u;
    const uint32_t* base = reinterpret_cast<const uint32_t*>(A.codes + wp.code_off);
    auto code_at = [&](uint32_t i, uint32_t j) -> uint32_t {                 // 1 <= i <= m, 1 <= j <= n
        const uint32_t x = i - 1u, band = x / (32u * R), L = (x / R) & 31u, r = x % R, q = j + L;
        const uint64_t chunk = (((uint64_t)band * R + r) * nblk + (q >> 5)) * 32u + L;
        return (__ldg(base + chunk * 4u + ((q & 31u) >> 3)) >> (4u * (q & 7u))) & 15u;
    };
    WarpOpsSink sink(A.ops + A.ops_off[wp.pair], lane == 0);
    uint32_t i = wp.m, j = wp.n, nops = 0;
    uint32_t state = (i && j) ? (code_at(i, j) & 3u) : 0u;                   // hw3.cpp:86-98
    while (i > 0u || j > 0u) {
        if (j == 0u) { sink.put_run(OP_D, i); nops += i; break; }            // column 0 is all F (traceF[i][0], hw3.cpp:44-45)
        if (i == 0u) { sink.put_run(OP_I, j); nops += j; break; }            // row 0 is all E    (traceE[0][j], hw3.cpp:50-51)
        const uint32_t k = (uint32_t)lane + 1u;
        const uint32_t ci = state == 2u ? i : i - min(k, i), cj = state == 1u ? j : j - min(k, j);
        const bool valid = ci >= 1u && cj >= 1u && (state == 2u || i > (uint32_t)lane) && (state == 1u || j > (uint32_t)lane);
        const uint32_t c = valid ? code_at(ci, cj) : 0u;
        const bool cont = valid && (state == 0u ? (c & 3u) == 0u : (state == 1u ? (c & 4u) != 0u : (c & 8u) != 0u));
        const uint32_t okm = __ballot_sync(0xFFFFFFFFu, cont);
        const uint32_t lead = okm == 0xFFFFFFFFu ? 32u : (uint32_t)(__ffs(~okm) - 1);
        const uint32_t steps = min(lead + 1u, 32u);                          // the current cell's own move + the confirmed ones
        const uint32_t op = state == 0u ? OP_M : (state == 1u ? OP_D : OP_I);
        sink.put_run(op, steps); nops += steps;
        if (state != 2u) i -= steps;
        if (state != 1u) j -= steps;
        if (lead < 32u) {                                                    // lane `lead` saw where the run ends
            const uint32_t v_t = __shfl_sync(0xFFFFFFFFu, (uint32_t)valid, (int)lead);
            const uint32_t c_t = __shfl_sync(0xFFFFFFFFu, c, (int)lead);
            if (v_t) state = state == 0u ? (c_t & 3u) : 0u;                  // traceV names the next state; a gap that stops extending was opened from V
            // !v_t: a border was reached (i == 0 or j == 0), handled at the top of the loop
        }
    }
    sink.flush();
    if (lane == 0) A.n_ops[wp.pair] = nops;
}

} // namespace b2a
