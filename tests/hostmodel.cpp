// hostmodel.cpp -- TEST INFRASTRUCTURE.  CPU model of the short16 HBM record.
//
// Builds the record the fill kernel is specified to write (b2a_format.h) from a
// plainly computed DP matrix, then runs the SAME traceback walkers the CUDA
// traceback kernel runs (they live in b2a_format.h).  tests/test_format_model.py
// checks the walkers' output against the oracle on CPU, and the GPU tests diff
// the fill kernel's real buffers against hm_encode()'s expectation.
// Never linked into the product library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../bioinformatics-algorithms_b200/csrc/b2a_format.h"
#include "../bioinformatics-algorithms_b200/csrc/fasta_hw2.h"

using namespace b2a;
static int g_opt = 3;
extern "C" void hm_set_opt(int o) { g_opt = o; }

namespace {

// biased H of the extended problem: 32R rows (junk rows use byte `junk` as pattern symbol), n columns
std::vector<int32_t> dp_matrix(int mode, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int R,
                               int match, int mismatch, int gap, int bias, uint8_t junk) {
    const uint32_t rows = 32u * (uint32_t)R;
    const size_t W = (size_t)n + 1;
    std::vector<int32_t> H((rows + 1) * W);
    for (uint32_t j = 0; j <= n; ++j) H[j] = mode == 0 ? bias + (int)j * gap : 0;
    for (uint32_t i = 1; i <= rows; ++i) {
        H[i * W] = mode == 0 ? bias + (int)i * gap : 0;
        const uint8_t pc = i <= m ? p[i - 1] : junk;
        for (uint32_t j = 1; j <= n; ++j) {
            int v = H[(i - 1) * W + j - 1] + (pc == t[j - 1] ? match : mismatch);
            const int l = H[i * W + j - 1] + gap, u = H[(i - 1) * W + j] + gap;
            if (l > v) v = l;
            if (u > v) v = u;
            if (mode != 0 && v < 0) v = 0;
            H[i * W + j] = v;
        }
    }
    return H;
}

template <int K>
int encode(int mode, const uint8_t* pa, const uint8_t* pb, uint32_t m, const uint8_t* ta, const uint8_t* tb, uint32_t n,
           int R, int match, int mismatch, int gap, int bias, uint8_t junk, Chunk* codes, uint32_t* rowbest) {
    constexpr int F = Geo<K>::F, CS = Geo<K>::CS;
    const uint32_t NC = num_chunks(n, CS);
    const size_t W = (size_t)n + 1;
    auto HA = dp_matrix(mode, pa, m, ta, n, R, match, mismatch, gap, bias, junk);
    auto HB = dp_matrix(mode, pb, m, tb, n, R, match, mismatch, gap, bias, junk);
    for (uint32_t L = 0; L < 32; ++L)
        for (int r = 0; r < R; ++r) {
            const uint32_t i = L * (uint32_t)R + (uint32_t)r + 1;
            auto P = [&](int64_t q) -> uint32_t {           // packed H after step q (frozen outside 1..n)
                int64_t j = q - (int64_t)L;
                if (j < 0) j = 0;
                if (j > (int64_t)n) j = n;
                const int32_t a = HA[i * W + (size_t)j], b = HB[i * W + (size_t)j];
                if (a < 0 || a > 32767 || b < 0 || b > 32767) return 0xFFFFFFFFu;   // flagged below
                return (uint32_t)a | ((uint32_t)b << 16);
            };
            int bestA = 0, bestB = 0;
            for (uint32_t j = 1; j <= n; ++j) {
                if (HA[i * W + j] > bestA) bestA = HA[i * W + j];
                if (HB[i * W + j] > bestB) bestB = HB[i * W + j];
            }
            if (rowbest) rowbest[(uint32_t)r * 32u + L] = (uint32_t)bestA | ((uint32_t)bestB << 16);
            for (uint32_t c = 0; c < NC; ++c) {
                uint32_t w[3];
                for (int wi = 0; wi < 3; ++wi) {
                    const int64_t q0 = (int64_t)c * CS + (int64_t)wi * F;
                    uint32_t Pv[F + 1];
                    for (int t = 0; t <= F; ++t) {
                        Pv[t] = P(q0 - 1 + t);
                        if (Pv[t] == 0xFFFFFFFFu) return -2;                     // value outside [0, 32767]
                    }
                    w[wi] = encode_word<Short16<K>>(Pv, gap);
                    // cross-check the ring-arithmetic word against direct field packing
                    uint32_t direct = 0;
                    for (int t = 0; t < F; ++t) {
                        const int dlo = (int)(Pv[t + 1] & 0xFFFF) - (int)(Pv[t] & 0xFFFF) - gap;
                        const int dhi = (int)(Pv[t + 1] >> 16) - (int)(Pv[t] >> 16) - gap;
                        if (dlo < 0 || dlo > (int)Geo<K>::MASK || dhi < 0 || dhi > (int)Geo<K>::MASK) return -3;   // lemma violated
                        direct |= ((uint32_t)dlo << (K * (F - 1 - t))) | ((uint32_t)dhi << (16 + K * (F - 1 - t)));
                    }
                    if (direct != w[wi]) return -4;
                }
                codes[((size_t)c * 32u + L) * (uint32_t)R + (uint32_t)r] = Chunk{w[0], w[1], w[2], P((int64_t)c * CS + CS - 1)};   // short16: row-major inside a chunk column
            }
        }
    return 0;
}

struct HostLoader {
    const Chunk* base;
    Chunk operator()(uint64_t idx) const { return base[idx]; }
    void prefetch(uint64_t) const {}
};

template <int K>
int run(int mode, const uint8_t* pa, const uint8_t* pb, uint32_t m, const uint8_t* ta, const uint8_t* tb, uint32_t n,
        int R, int match, int mismatch, int gap, int bias, uint8_t junk, PairResult* res, uint32_t* opsA, uint32_t* opsB,
        Chunk* codes_out, uint32_t* rowbest_out) {
    const uint32_t NC = num_chunks(n, Geo<K>::CS);
    std::vector<Chunk> codes((size_t)R * NC * 32);
    std::vector<uint32_t> rowbest((size_t)R * 32);
    int rc = encode<K>(mode, pa, pb, m, ta, tb, n, R, match, mismatch, gap, bias, junk, codes.data(), rowbest.data());
    if (rc) return rc;
    if (codes_out) std::memcpy(codes_out, codes.data(), codes.size() * sizeof(Chunk));
    if (rowbest_out) std::memcpy(rowbest_out, rowbest.data(), rowbest.size() * sizeof(uint32_t));
    for (int half = 0; half < 2; ++half) {
        PairView v{codes.data(), rowbest.data(), half ? pb : pa, half ? tb : ta, m, n, NC, R, half, match, mismatch, gap, bias, g_opt, short16_rmagic(R)};
        OpsSink sink(half ? opsB : opsA);
        std::memset(&res[half], 0, sizeof(PairResult));
        if (mode == 0) walk_global<Short16<K>>(v, HostLoader{codes.data()}, sink, res[half]);
        else walk_local<Short16<K>>(v, HostLoader{codes.data()}, sink, res[half]);
        sink.flush();
        res[half].path = 1;
    }
    return 0;
}


// ---- wide32 record model: one pair, bands of 128 rows, int32, Wide32<K> chunks ----
template <int K>
int run_wide(int mode, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int match, int mismatch, int gap,
             PairResult* res, uint32_t* ops) {
    using FM = Wide32<K>;
    constexpr int F = FM::F, CS = FM::CS, R = WIDE_R;
    const uint32_t NC = num_chunks(n, CS, FM::SKEW);
    const uint32_t nbands = (m + 32u * R - 1u) / (32u * R);
    const size_t W = (size_t)n + 1;
    // plain DP, unbiased; rows beyond m never match anything (as in the kernel)
    const uint32_t rows = nbands * 32u * R;
    std::vector<int32_t> H((size_t)(rows + 1) * W);
    for (uint32_t j = 0; j <= n; ++j) H[j] = mode == 0 ? (int)j * gap : 0;
    for (uint32_t i = 1; i <= rows; ++i) {
        H[i * W] = mode == 0 ? (int)i * gap : 0;
        for (uint32_t j = 1; j <= n; ++j) {
            const bool eq = i <= m && p[i - 1] == t[j - 1];
            int v = H[(i - 1) * W + j - 1] + (eq ? match : mismatch);
            const int l = H[i * W + j - 1] + gap, u = H[(i - 1) * W + j] + gap;
            if (l > v) v = l;
            if (u > v) v = u;
            if (mode != 0 && v < 0) v = 0;
            H[i * W + j] = v;
        }
    }
    std::vector<Chunk> codes((size_t)nbands * R * NC * 32);
    std::vector<uint32_t> rowbest((size_t)nbands * R * 32);
    for (uint32_t band = 0; band < nbands; ++band)
        for (uint32_t L = 0; L < 32; ++L)
            for (int r = 0; r < R; ++r) {
                const uint32_t i = band * 32u * R + L * R + (uint32_t)r + 1;
                auto P = [&](int64_t q) -> uint32_t {
                    int64_t j = q - (int64_t)L * FM::SKEW;
                    if (j < 0) j = 0;
                    if (j > (int64_t)n) j = n;
                    return (uint32_t)H[i * W + (size_t)j];
                };
                int best = 0;
                for (uint32_t j = 1; j <= n; ++j) if (H[i * W + j] > best) best = H[i * W + j];
                rowbest[((size_t)band * R + r) * 32u + L] = (uint32_t)best;
                for (uint32_t c = 0; c < NC; ++c) {
                    uint32_t w[2];
                    for (int wi = 0; wi < 2; ++wi) {
                        const int64_t q0 = (int64_t)c * CS + (int64_t)wi * F;
                        uint32_t Pv[F + 1];
                        for (int tt = 0; tt <= F; ++tt) Pv[tt] = P(q0 - 1 + tt);
                        w[wi] = encode_word<FM>(Pv, gap);
                        if (K < 32) {
                            uint32_t direct = 0;
                            for (int tt = 0; tt < F; ++tt) {
                                const int64_t d = (int64_t)(int32_t)Pv[tt + 1] - (int32_t)Pv[tt] - gap;
                                if (d < 0 || d > (int64_t)FM::MASK) return -3;
                                direct |= (uint32_t)d << ((K & 31) * (F - 1 - tt));
                            }
                            if (direct != w[wi]) return -4;
                        }
                    }
                    codes[(((size_t)band * R + r) * NC + c) * 32u + L] = Chunk{w[0], w[1], 0u, P((int64_t)c * CS + CS - 1)};
                }
            }
    PairView v{codes.data(), rowbest.data(), p, t, m, n, NC, R, 0, match, mismatch, gap, 0, g_opt, 0};
    OpsSink sink(ops);
    std::memset(res, 0, sizeof(PairResult));
    if (mode == 0) walk_global<FM>(v, HostLoader{codes.data()}, sink, *res);
    else walk_local<FM>(v, HostLoader{codes.data()}, sink, *res);
    sink.flush();
    res->path = 2;
    return 0;
}

} // namespace

extern "C" {

// bin/hw2's FASTA reader with a forced thread count (min_parallel_bytes = 0): count of records, bytes and offsets into caller buffers;
// returns the record count, -1 if the file cannot be opened, -2 if a buffer is too small
int64_t hm_load_fasta(const char* path, unsigned threads, uint8_t* bytes, uint64_t bytes_cap, uint64_t* off, uint64_t off_cap) {
    b2a_cli::FastaBatch fb;
    if (!b2a_cli::load_fasta(path, fb, 0, threads)) return -1;
    if (fb.size > bytes_cap || fb.off.size() > off_cap) return -2;
    std::memcpy(bytes, fb.data, fb.size);
    std::memcpy(off, fb.off.data(), fb.off.size() * sizeof(uint64_t));
    return (int64_t)fb.count();
}

// number of chunks in one pair-pair record
uint64_t hm_record_chunks(int K, int R, uint32_t n) {
    const int CS = K == 2 ? Geo<2>::CS : K == 4 ? Geo<4>::CS : Geo<8>::CS;
    return (uint64_t)R * num_chunks(n, CS) * 32u;
}

// exhaustive check of the multiply-shift row -> lane division used by the short16 walkers
int hm_check_rmagic() {
    for (int R = 1; R <= SHORT16_MAX_R; ++R)
        for (uint32_t x = 0; x < 32u * R; ++x)
            if (((x * (uint32_t)short16_rmagic(R)) >> 16) != x / (uint32_t)R) return 0;
    return 1;
}
int hm_delta_bits(int match, int mismatch, int gap) { return delta_bits(match, mismatch, gap); }
int hm_delta_bits_wide(int match, int mismatch, int gap) { return delta_bits_wide(match, mismatch, gap); }

// Model one pair through the wide32 record (K <= 0: choose like the library does).
int hm_run_wide(int mode, int K, const uint8_t* p, uint32_t m, const uint8_t* t, uint32_t n, int match, int mismatch, int gap,
                PairResult* res, uint32_t* ops) {
    if (K <= 0) K = delta_bits_wide(match, mismatch, gap);
    switch (K) {
        case 2:  return run_wide<2>(mode, p, m, t, n, match, mismatch, gap, res, ops);
        case 4:  return run_wide<4>(mode, p, m, t, n, match, mismatch, gap, res, ops);
        case 8:  return run_wide<8>(mode, p, m, t, n, match, mismatch, gap, res, ops);
        case 16: return run_wide<16>(mode, p, m, t, n, match, mismatch, gap, res, ops);
        case 32: return run_wide<32>(mode, p, m, t, n, match, mismatch, gap, res, ops);
    }
    return -1;
}

// out[0..2] = K, R, bias; returns 1 if the short16 record can hold the pair-class
int hm_plan(int mode, uint32_t m, uint32_t n, int match, int mismatch, int gap, int* out) {
    Short16Plan pl{0, 0, 0};
    if (!short16_plan(mode, m, n, match, mismatch, gap, pl)) return 0;
    out[0] = pl.K; out[1] = pl.R; out[2] = pl.bias;
    return 1;
}

// Model one pair-pair end to end. res[2]; opsA/opsB hold ceil((m+n)/16) words each (may be null).
int hm_run_pairpair(int mode, int K, int R, const uint8_t* pa, const uint8_t* pb, uint32_t m, const uint8_t* ta,
                    const uint8_t* tb, uint32_t n, int match, int mismatch, int gap, int bias, uint8_t junk,
                    PairResult* res, uint32_t* opsA, uint32_t* opsB, Chunk* codes_out, uint32_t* rowbest_out) {
    switch (K) {
        case 2: return run<2>(mode, pa, pb, m, ta, tb, n, R, match, mismatch, gap, bias, junk, res, opsA, opsB, codes_out, rowbest_out);
        case 4: return run<4>(mode, pa, pb, m, ta, tb, n, R, match, mismatch, gap, bias, junk, res, opsA, opsB, codes_out, rowbest_out);
        case 8: return run<8>(mode, pa, pb, m, ta, tb, n, R, match, mismatch, gap, bias, junk, res, opsA, opsB, codes_out, rowbest_out);
    }
    return -1;
}

} // extern "C"
