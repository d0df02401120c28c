// b2a_format.h -- HBM layout of the short-pair ("short16") DP record and the
// tie-exact traceback that reads it.  Shared by the CUDA kernels (device) and
// by the host-side format model used in CPU tests (tests/hostmodel.cpp), so the
// decoder logic is exercised without a GPU.  No CUDA types in this header.
//
// What is stored (see DESIGN.md "short16 record"):
//   The fill kernel never writes H or direction letters.  Per DP row it writes
//   the HORIZONTAL DELTAS  D(i,j) = H(i,j) - H(i,j-1) - gap, which for a linear
//   gap always lie in [0, max(match,mismatch) - 2*gap]  (K = 2, 4 or 8 bits),
//   plus one absolute 16-bit anchor H per chunk.  The traceback rebuilds the
//   exact H of the three neighbours of each visited cell and then applies the
//   reference's own comparisons (hw2.cpp:145-153 for NW, hw2.cpp:214-222 for SW),
//   so ties break exactly as in the reference by construction.
//
// Geometry for one "pair-pair" (two pairs of identical shape m x n packed in the
// low/high 16-bit halves of every word):
//   lane L (0..31) owns rows i = L*R + r + 1, r = 0..R-1 (rows beyond m are junk);
//   at wavefront step q lane L is at column j = q - L (active iff 1 <= j <= n,
//   otherwise its H registers are frozen, which makes D = -gap, still in range);
//   a word holds F = 16/K steps per half, a chunk = {w0, w1, w2, anchor} = 16 B
//   covers CS = 3F steps; anchor = packed H after the chunk's last step.
//   chunk index of (r, c, L) = (r*NC + c)*32 + L  -> a warp stores 512 contiguous bytes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2A_HD __host__ __device__ __forceinline__
#else
#define B2A_HD inline
#endif

namespace b2a {

struct alignas(16) Chunk { uint32_t w0, w1, w2, anchor; };

// pair-pair descriptor: pairs a and b (b == a for an unpaired singleton) of shape m x n
struct PPDesc { uint32_t a, b, m, n; };

// one record per pair; mirrors b2a_result in include/b2align.h (static_assert in b2a_api.cu)
struct PairResult {
    int32_t score; uint32_t end_i, end_j, start_i, start_j; int32_t overlap; uint32_t n_ops; uint32_t path;
};

enum : uint32_t { OP_M = 0, OP_D = 1, OP_I = 2 };

template <int K> struct Geo {
    static_assert(K == 2 || K == 4 || K == 8, "delta width");
    static constexpr int F = 16 / K;        // steps per 16-bit half-word
    static constexpr int CS = 3 * F;        // steps per 16-byte chunk
    static constexpr uint32_t MASK = (1u << K) - 1u;
};

B2A_HD uint32_t num_chunks(uint32_t n, int CS) { return (n + 32u + (uint32_t)CS - 1u) / (uint32_t)CS; }

// smallest supported K for a scoring scheme, 0 if the short16 record cannot hold it
B2A_HD int delta_bits(int match, int mismatch, int gap) {
    int smax = match > mismatch ? match : mismatch;
    if (gap > 0 || smax < 0) return 0;
    long W = (long)smax - 2L * gap + 1;       // number of distinct D values
    if (W <= 4) return 2;
    if (W <= 16) return 4;
    if (W <= 256) return 8;
    return 0;
}


// Can the s16x2 record hold an m x n pair-class under this scoring?  Chooses rows-per-lane R,
// the delta width K and the bias that keeps every stored H of the (junk-extended) matrix in
// [0, 32767] (needed both for the s16 arithmetic and for the ring-exact word encoding).
struct Short16Plan { int K, R, bias; };
constexpr int SHORT16_MAX_R = 8;          // rows per lane -> patterns up to 256 bases
B2A_HD bool short16_plan(int mode, uint32_t m, uint32_t n, int match, int mismatch, int gap, Short16Plan& pl) {
    if (m == 0 || n == 0 || m > 32u * SHORT16_MAX_R || n > 60000u) return false;
    pl.K = delta_bits(match, mismatch, gap);
    if (pl.K == 0) return false;
    if (match > 127 || match < -128 || mismatch > 127 || mismatch < -128) return false;   // PRMT score tables are int8
    pl.R = (int)((m + 31u) / 32u);
    const long rows = 32L * pl.R;
    const long smax = match > mismatch ? match : mismatch, smin = match < mismatch ? match : mismatch;
    const long margin = (long)(-gap) + (smin < 0 ? -smin : 0) + 1;
    const long top = (rows < (long)n ? rows : (long)n) * (smax > 0 ? smax : 0) + margin;
    long bias = 0;
    if (mode == 0) bias = (rows + (long)n) * (long)(-gap) + margin;
    if (bias + top > 32767) return false;
    pl.bias = (int)bias;
    return true;
}

B2A_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// sum of the K-bit fields of x (x < 2^48)
template <int K> B2A_HD int field_sum(uint64_t x);
template <> B2A_HD int field_sum<2>(uint64_t x) {
    return popc64(x & 0x5555555555555555ull) + 2 * popc64(x & 0xAAAAAAAAAAAAAAAAull);
}
template <> B2A_HD int field_sum<4>(uint64_t x) {
    uint64_t b = (x & 0x0F0F0F0F0F0F0F0Full) + ((x >> 4) & 0x0F0F0F0F0F0F0F0Full);   // byte sums <= 30
    return (int)((b * 0x0101010101010101ull) >> 56);
}
template <> B2A_HD int field_sum<8>(uint64_t x) {
    uint64_t h = (x & 0x00FF00FF00FF00FFull) + ((x >> 8) & 0x00FF00FF00FF00FFull);   // 16-bit sums <= 510
    return (int)(((h * 0x0001000100010001ull) >> 48) & 0xFFFF);
}

// Read-only view of one pair inside its pair-pair's record.
struct PairView {
    const Chunk*    codes;     // this pair-pair's chunk block
    const uint32_t* rowbest;   // [R][32] packed per-row maxima (local mode), else unused
    const uint8_t*  p;         // pattern bytes (rows)
    const uint8_t*  t;         // text bytes (columns)
    uint32_t m, n, NC;
    int R, half;               // half: 0 = low 16 bits, 1 = high
    int match, mismatch, gap, bias;
};

template <int K>
B2A_HD uint64_t chunk_bits(const Chunk& ch, int half) {
    const int sh = half * 16;
    return ((uint64_t)((ch.w0 >> sh) & 0xFFFFu) << 32) | ((uint64_t)((ch.w1 >> sh) & 0xFFFFu) << 16) |
           (uint64_t)((ch.w2 >> sh) & 0xFFFFu);
}

// H (biased) and D of the cell computed at in-chunk step `rem`
template <int K>
B2A_HD void decode_step(const Chunk& ch, int half, int gap, int rem, int& H, int& D) {
    constexpr int CS = Geo<K>::CS;
    const uint64_t X = chunk_bits<K>(ch, half);
    const int anchor = (int)((ch.anchor >> (half * 16)) & 0xFFFFu);
    const int off = K * (CS - 1 - rem);
    D = (int)((X >> off) & Geo<K>::MASK);
    const uint64_t low = X & ((1ull << off) - 1ull);
    H = anchor - (CS - 1 - rem) * gap - field_sum<K>(low);
}

// Two-entry cache of chunks in registers: the walk alternates between the current row and the row above.
template <class Loader>
struct ChunkCache {
    Loader ld;
    uint32_t i0, i1;
    Chunk c0, c1;
    B2A_HD explicit ChunkCache(Loader l) : ld(l), i0(0xFFFFFFFFu), i1(0xFFFFFFFFu) { c0 = Chunk{0, 0, 0, 0}; c1 = c0; }
    B2A_HD Chunk get(uint32_t idx) {
        if (idx == i0) return c0;
        if (idx == i1) return c1;
        c1 = c0; i1 = i0;
        c0 = ld(idx); i0 = idx;
        return c0;
    }
};

template <int K, class Loader>
B2A_HD void cell(const PairView& v, ChunkCache<Loader>& cc, uint32_t i, uint32_t j, int& H, int& D) {
    constexpr int CS = Geo<K>::CS;
    const uint32_t L = (i - 1u) / (uint32_t)v.R, r = (i - 1u) - L * (uint32_t)v.R;
    const uint32_t q = j + L, c = q / (uint32_t)CS, rem = q - c * (uint32_t)CS;
    const Chunk ch = cc.get((r * v.NC + c) * 32u + L);
    decode_step<K>(ch, v.half, v.gap, (int)rem, H, D);
}

// ---- Needleman-Wunsch traceback, hw2.cpp:158-181 on reconstructed H; directions per hw2.cpp:145-153 ----
template <int K, class Loader, class Sink>
B2A_HD void walk_global(const PairView& v, Loader ld, Sink& sink, PairResult& res) {
    ChunkCache<Loader> cc(ld);
    uint32_t i = v.m, j = v.n, nops = 0;
    int cur = 0, best = 0, H, D;
    if (i == 0 || j == 0) H = v.bias + (int)(i + j) * v.gap;       // border, hw2.cpp:125-136
    else cell<K>(v, cc, i, j, H, D);
    res.score = H - v.bias;                                         // hw2.cpp:186
    while (i > 0 || j > 0) {
        uint32_t op;
        if (i == 0) op = OP_I;                                      // row 0 holds 'l' (hw2.cpp:134)
        else if (j == 0) op = OP_D;                                 // column 0 holds 'u' (hw2.cpp:128)
        else {
            cell<K>(v, cc, i, j, H, D);
            const int Hl = H - D - v.gap;                           // H(i, j-1); frozen lanes make this the border at j == 1
            int Hu, Hd;
            if (i == 1) { Hu = v.bias + (int)j * v.gap; Hd = v.bias + (int)(j - 1) * v.gap; }
            else { int Du; cell<K>(v, cc, i - 1, j, Hu, Du); Hd = Hu - Du - v.gap; }
            const bool eq = v.p[i - 1] == v.t[j - 1];
            int val = Hd + (eq ? v.match : v.mismatch);             // hw2.cpp:142
            op = OP_M;                                              // hw2.cpp:145
            if (Hl + v.gap > val) { val = Hl + v.gap; op = OP_I; }  // hw2.cpp:146-149 'l'
            if (Hu + v.gap > val) { op = OP_D; }                    // hw2.cpp:150-153 'u'
        }
        if (op == OP_M) {
            const uint8_t pc = v.p[i - 1];
            if (pc == v.t[j - 1] && pc != (uint8_t)'-') { if (++cur > best) best = cur; } else cur = 0;   // hw2.cpp:267-278
            --i; --j;
        } else if (op == OP_D) { cur = 0; --i; }
        else { cur = 0; --j; }
        sink.put(op);
        ++nops;
    }
    res.end_i = v.m; res.end_j = v.n; res.start_i = 0; res.start_j = 0;
    res.overlap = best; res.n_ops = nops;
}

// ---- Smith-Waterman: first row-major arg-max (hw2.cpp:225-229) + traceback hw2.cpp:239-257 ----
template <int K, class Loader, class Sink>
B2A_HD void walk_local(const PairView& v, Loader ld, Sink& sink, PairResult& res) {
    constexpr int CS = Geo<K>::CS;
    ChunkCache<Loader> cc(ld);
    const int sh = v.half * 16;
    // per-row maxima written by the fill kernel -> global max M and the first row that attains it
    int M = 0; uint32_t bi = 0;
    for (uint32_t i = 1; i <= v.m; ++i) {
        const uint32_t L = (i - 1u) / (uint32_t)v.R, r = (i - 1u) - L * (uint32_t)v.R;
        const int rb = (int)((v.rowbest[r * 32u + L] >> sh) & 0xFFFFu);
        if (rb > M) { M = rb; bi = i; }
    }
    res.score = M; res.overlap = 0; res.n_ops = 0;
    if (M == 0) { res.end_i = res.end_j = res.start_i = res.start_j = 0; return; }   // hw2.cpp:202-203: best cell stays (0,0)
    // first column of row bi whose H equals M: rebuild the row left to right from its deltas (H(bi,0) = 0)
    uint32_t bj = 0;
    {
        const uint32_t L = (bi - 1u) / (uint32_t)v.R, r = (bi - 1u) - L * (uint32_t)v.R;
        int run = 0;
        uint32_t q = L + 1u;
        while (q <= L + v.n && bj == 0) {
            const uint32_t c = q / (uint32_t)CS;
            const Chunk ch = ld((r * v.NC + c) * 32u + L);
            const uint64_t X = chunk_bits<K>(ch, v.half);
            uint32_t rem = q - c * (uint32_t)CS;
            for (; rem < (uint32_t)CS && q <= L + v.n; ++rem, ++q) {
                run += (int)((X >> (K * (CS - 1 - (int)rem))) & Geo<K>::MASK) + v.gap;
                if (run == M) { bj = q - L; break; }
            }
        }
    }
    uint32_t i = bi, j = bj, nops = 0;
    int cur = 0, best = 0, H = M;
    while (i > 0 && j > 0 && H != 0) {                               // hw2.cpp:239
        int Hc, D;
        cell<K>(v, cc, i, j, Hc, D);
        const int Hl = Hc - D - v.gap;
        int Hu = 0, Hd = 0;
        if (i > 1) { int Du; cell<K>(v, cc, i - 1, j, Hu, Du); Hd = Hu - Du - v.gap; }
        const uint8_t pc = v.p[i - 1];
        const bool eq = pc == v.t[j - 1];
        uint32_t op;                                                 // hw2.cpp:214-222 (H != 0 here)
        if (H == Hd + (eq ? v.match : v.mismatch)) op = OP_M;
        else if (H == Hu + v.gap) op = OP_D;
        else op = OP_I;
        if (op == OP_M) {
            if (eq && pc != (uint8_t)'-') { if (++cur > best) best = cur; } else cur = 0;
            --i; --j; H = Hd;
        } else if (op == OP_D) { cur = 0; --i; H = Hu; }
        else { cur = 0; --j; H = Hl; }
        sink.put(op);
        ++nops;
    }
    res.end_i = bi; res.end_j = bj; res.start_i = i; res.start_j = j;
    res.overlap = best; res.n_ops = nops;
}

// 2-bit op writer: op t of a pair lands in bits 2*(t%16) of word t/16 (traceback order)
struct OpsSink {
    uint32_t* out;      // may be null: ops are counted but not stored
    uint32_t word, fill, pos;
    B2A_HD explicit OpsSink(uint32_t* o) : out(o), word(0), fill(0), pos(0) {}
    B2A_HD void put(uint32_t op) {
        word |= op << (2u * fill);
        if (++fill == 16u) { if (out) out[pos] = word; ++pos; word = 0; fill = 0; }
    }
    B2A_HD void flush() { if (fill && out) out[pos] = word; }
};

// Word the fill kernel emits for F consecutive steps of one row (both halves at once):
//   sum_t B^(F-1-t) * (P_t - P_{t-1} - g32)  (mod 2^32),  B = 2^K, g32 = gap * 65537,
// evaluated as  P_{F-1} + (B-1) * S - B^(F-1) * P_{-1} - g32 * (B^F - 1)/(B - 1)  with the Horner sum
// S = sum_{t<F-1} B^(F-2-t) P_t.  Exact per 16-bit half because every stored P has both halves in
// [0, 32767] and every true field lies in [0, B-1].
template <int K>
B2A_HD uint32_t encode_word(const uint32_t* P /* P[0] = P_{-1}, P[1..F] = P_0..P_{F-1} */, int gap) {
    constexpr int F = Geo<K>::F;
    const uint32_t B = 1u << K, g32 = (uint32_t)gap * 65537u;
    uint32_t S = 0;
    for (int t = 0; t < F - 1; ++t) S = S * B + P[t + 1];
    uint32_t Bpow = 1; for (int t = 0; t < F - 1; ++t) Bpow *= B;          // B^(F-1)
    uint32_t geo = 0; for (int t = 0; t < F; ++t) geo = geo * B + 1u;      // (B^F - 1)/(B - 1)
    return P[F] + (B - 1u) * S - Bpow * P[0] - g32 * geo;
}

} // namespace b2a
