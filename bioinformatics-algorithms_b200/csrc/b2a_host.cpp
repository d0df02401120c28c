// b2a_host.cpp -- host-only pieces of the C ABI: result formatting and batch selection.
// (prepareCigarString hw2.cpp:59-78, prepareMDZString hw2.cpp:80-116, selection hw2.cpp:326-357.)
// Works on the op list in traceback order plus the raw sequences; no aligned strings are built.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2align.h"

namespace {
struct Out {
    char* p; uint64_t cap, len; bool ok;
    void ch(char c) { if (len + 1 < cap) p[len] = c; else ok = false; ++len; }
    void num(uint64_t v) { char b[24]; int k = std::snprintf(b, sizeof b, "%llu", (unsigned long long)v); for (int i = 0; i < k; ++i) ch(b[i]); }
    int64_t done() { if (cap) p[len < cap ? len : cap - 1] = 0; return ok ? (int64_t)len : (int64_t)B2A_ERR_ARG; }
};
} // namespace

extern "C" {

// Run-length encode the alignment columns in alignment order (the list is stored end -> start).
int64_t b2a_render_cigar(const char* ops, uint64_t n_ops, char* out, uint64_t cap)
{
    if ((!ops && n_ops) || !out) return B2A_ERR_ARG;
    Out o{out, cap, 0, true};
    uint64_t k = n_ops;
    while (k > 0) {
        const char op = ops[k - 1];
        uint64_t run = 0;
        while (k > 0 && ops[k - 1] == op) { ++run; --k; }
        o.num(run); o.ch(op);
    }
    return o.done();
}

// MD:Z exactly as the reference emits it: a count of matching 'M' columns is always written before a
// mismatching reference base, before '^'+deleted pattern bases of a whole 'D' run, and at the end
// (so "0" appears freely); 'I' columns advance the text without touching the count.
int64_t b2a_render_mdz(const char* ops, uint64_t n_ops, const uint8_t* pattern, const uint8_t* text,
                       uint32_t start_i, uint32_t start_j, char* out, uint64_t cap)
{
    if ((!ops && n_ops) || !out || ((!pattern || !text) && n_ops)) return B2A_ERR_ARG;
    Out o{out, cap, 0, true};
    uint64_t pi = start_i, tj = start_j, same = 0, k = n_ops;
    while (k > 0) {
        const char op = ops[k - 1];
        if (op == 'M') {
            if (pattern[pi] == text[tj]) ++same;
            else { o.num(same); o.ch((char)text[tj]); same = 0; }
            ++pi; ++tj; --k;
        } else if (op == 'D') {
            o.num(same); o.ch('^'); same = 0;
            while (k > 0 && ops[k - 1] == 'D') { o.ch((char)pattern[pi]); ++pi; --k; }
        } else { ++tj; --k; }
    }
    o.num(same);
    return o.done();
}

int64_t b2a_select_best(int32_t mode, const b2a_result* results, uint64_t n_pairs)
{
    if (!results && n_pairs) return B2A_ERR_ARG;
    int64_t best_key = -1000000, best = -1;            // hw2.cpp:326
    for (uint64_t k = 0; k < n_pairs; ++k) {
        const int64_t key = mode == B2A_MODE_GLOBAL ? results[k].overlap : results[k].score;
        if (key > best_key) { best_key = key; best = (int64_t)k; }
    }
    return best;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// hw4's tree stage (hw4/hw4.cpp:154-228): UPGMA over the pair distances, Newick text.
// Bit-exact contract with the reference: the same double arithmetic in the same order
// (size-weighted average (d_a * size_a + d_b * size_b) / (size_a + size_b), height = d_min / 2), the
// first strict minimum in row-major order over the CURRENT cluster order, the merged cluster appended
// behind the survivors, branch lengths printed with std::to_string (6 decimals).
// Kept as an index list over one full matrix instead of rebuilding matrices per merge.
// ---------------------------------------------------------------------------------------------

extern "C" int64_t b2a_upgma_newick(const int32_t* pair_dist, uint32_t n_seqs, const char* const* names, char* out, uint64_t cap)
{
    if (!out || cap == 0 || (n_seqs && !names) || (n_seqs > 1 && !pair_dist)) return B2A_ERR_ARG;
    if (n_seqs == 0) return B2A_ERR_ARG;                       // the reference indexes clusters[0] unconditionally
    const size_t total = 2 * (size_t)n_seqs - 1;               // leaves + internal nodes
    std::vector<double> D(total * total, 0.0);                 // distance between any two nodes ever alive
    std::vector<int> size(total, 1);
    std::vector<double> height(total, 0.0);
    std::vector<std::string> label(total);
    std::vector<size_t> alive;                                 // current cluster order (hw4.cpp:198-206)
    for (uint32_t i = 0; i < n_seqs; ++i) { label[i] = names[i] ? names[i] : ""; alive.push_back(i); }
    size_t p = 0;
    for (uint32_t i = 0; i < n_seqs; ++i)
        for (uint32_t j = i + 1; j < n_seqs; ++j, ++p) D[i * total + j] = D[j * total + i] = (double)pair_dist[p];
    size_t next = n_seqs;
    while (alive.size() > 1) {
        double dmin = std::numeric_limits<double>::infinity();
        size_t ai = 0, aj = 0;
        for (size_t x = 0; x < alive.size(); ++x)              // hw4.cpp:167-175: first strict minimum, row-major
            for (size_t y = x + 1; y < alive.size(); ++y) {
                const double d = D[alive[x] * total + alive[y]];
                if (d < dmin) { dmin = d; ai = x; aj = y; }
            }
        const size_t a = alive[ai], b = alive[aj], c = next++;
        size[c] = size[a] + size[b];
        height[c] = dmin / 2.0;
        label[c] = "(" + label[a] + ":" + std::to_string(std::fabs(height[c] - height[a])) + "," +
                   label[b] + ":" + std::to_string(std::fabs(height[c] - height[b])) + ")";
        std::vector<size_t> keep;
        for (size_t x = 0; x < alive.size(); ++x) if (x != ai && x != aj) keep.push_back(alive[x]);
        for (size_t k : keep) {                                // hw4.cpp:222: distance of the merged cluster to every survivor
            const double d = (D[a * total + k] * size[a] + D[b * total + k] * size[b]) / size[c];
            D[c * total + k] = D[k * total + c] = d;
        }
        keep.push_back(c);
        alive.swap(keep);
    }
    const std::string tree = label[alive[0]] + ":0.0;";
    if (tree.size() + 1 > cap) return B2A_ERR_ARG;
    std::memcpy(out, tree.data(), tree.size());
    out[tree.size()] = 0;
    return (int64_t)tree.size();
}

// ---------------------------------------------------------------------------------------------
// hw3's assembly stage (hw3.cpp:253-357): centre-star merge of the pairwise alignments + PHYLIP text.
// Works on op lists (traceback order; 'M' centre base over other base, 'D' centre base over '-',
// 'I' '-' over other base) instead of aligned strings: the gap pattern of pair i is the number of 'I'
// columns in front of every centre base (hw3.cpp:270-277), the merged pattern is the per-position
// maximum (hw3.cpp:284-289), and every row is written block by block: pair i's own inserted bases
// first, then padding '-' up to the merged width, then its column for the centre base -- which is what
// the reference's character walk (hw3.cpp:303-325) produces for centre sequences without a literal '-'.
// ---------------------------------------------------------------------------------------------
extern "C" int64_t b2a_center_star_phylip(uint32_t n_seqs, uint32_t centre, const char* const* names,
                                          const uint8_t* const* seqs, const uint64_t* seq_len,
                                          const char* const* ops, const uint64_t* n_ops, char* out, uint64_t cap)
{
    if (!out || cap == 0 || n_seqs == 0 || centre >= n_seqs || !names || !seqs || !seq_len || (n_seqs > 1 && (!ops || !n_ops))) return B2A_ERR_ARG;
    const uint64_t L = seq_len[centre];
    std::vector<std::vector<uint32_t>> gaps(n_seqs);                 // per pair: 'I' columns in front of centre position k (k = 0..L)
    std::vector<uint32_t> merged(L + 1, 0);
    for (uint32_t i = 0; i < n_seqs; ++i) {
        if (i == centre) continue;
        gaps[i].assign(L + 1, 0);
        uint64_t pos = 0;
        for (uint64_t t = n_ops[i]; t-- > 0;) {                      // alignment order = reverse traceback order
            const char op = ops[i][t];
            if (op == 'I') { if (pos > L) return B2A_ERR_ARG; ++gaps[i][pos]; }
            else if (op == 'M' || op == 'D') ++pos;
            else return B2A_ERR_ARG;
        }
        if (pos != L) return B2A_ERR_ARG;                            // the ops must consume the whole centre
        for (uint64_t k = 0; k <= L; ++k) merged[k] = std::max(merged[k], gaps[i][k]);
    }
    std::vector<std::string> rows(n_seqs);
    for (uint64_t k = 0; k <= L; ++k) {                              // centre row, hw3.cpp:293-299
        rows[centre].append(merged[k], '-');
        if (k < L) rows[centre].push_back((char)seqs[centre][k]);
    }
    for (uint32_t i = 0; i < n_seqs; ++i) {
        if (i == centre) continue;
        std::string& r = rows[i];
        r.reserve(rows[centre].size());
        uint64_t k = 0, q = 0, own = 0;                              // centre position, position in sequence i, inserted bases seen in this block
        for (uint64_t t = n_ops[i]; t-- > 0;) {
            const char op = ops[i][t];
            if (op == 'I') { if (q >= seq_len[i]) return B2A_ERR_ARG; r.push_back((char)seqs[i][q++]); ++own; continue; }
            r.append(merged[k] - own, '-');                          // pad the gap block of centre position k
            if (op == 'M') { if (q >= seq_len[i]) return B2A_ERR_ARG; r.push_back((char)seqs[i][q++]); }
            else r.push_back('-');
            ++k; own = 0;
        }
        r.append(merged[L] - own, '-');                              // trailing block, hw3.cpp:320-324
        if (q != seq_len[i]) return B2A_ERR_ARG;
    }
    // hw3.cpp:330-331: the centre trades places with the first sequence; hw3.cpp:339-357: 10-char id, blocks of 10
    std::vector<uint32_t> order(n_seqs);
    for (uint32_t i = 0; i < n_seqs; ++i) order[i] = i;
    std::swap(order[0], order[centre]);
    std::string text = std::to_string(n_seqs) + " " + std::to_string(rows[centre].size()) + "\n";
    for (uint32_t o = 0; o < n_seqs; ++o) {
        const uint32_t i = order[o];
        std::string id = names[i] ? names[i] : "";
        if (id.size() > 10) id.resize(10); else id.append(10 - id.size(), ' ');
        text += id;
        const std::string& r = rows[i];
        for (size_t j = 0; j < r.size(); ++j) { if (j && j % 10 == 0) text.push_back(' '); text.push_back(r[j]); }
        text.push_back('\n');
    }
    if (text.size() + 1 > cap) return B2A_ERR_ARG;
    std::memcpy(out, text.data(), text.size());
    out[text.size()] = 0;
    return (int64_t)text.size();
}

// ---- compact sequence input (b2a_seq2, include/b2align.h): host-side packer / unpacker --------------------------------------------
// The reference holds sequences as std::string and compares raw bytes (hw2.cpp:142, :208); the 2-bit form is this engine's wire format
// for them, lossless through the exception list.
extern "C" int64_t b2a_seq2_pack(const uint8_t* seq, uint64_t n_bytes, const uint8_t alphabet[4],
                                 uint8_t* codes, uint64_t* exc_pos, uint8_t* exc_byte, uint64_t exc_cap)
{
    if ((!seq && n_bytes) || !alphabet) return B2A_ERR_ARG;
    if (exc_cap && (!exc_pos || !exc_byte)) return B2A_ERR_ARG;
    uint8_t lut[256];
    std::memset(lut, 0xFF, sizeof lut);
    for (int c = 3; c >= 0; --c) lut[alphabet[c]] = (uint8_t)c;          // a repeated alphabet byte takes its lowest code
    // slices of whole code bytes, one per thread; pass 1 counts the exceptions, pass 2 writes codes and exceptions at their offsets
    const uint64_t quads = (n_bytes + 3) / 4;
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (const char* e = std::getenv("B2A_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(1, quads / 65536));
    std::vector<uint64_t> cnt(nt + 1, 0);
    auto slice = [&](unsigned t, uint64_t& b0, uint64_t& b1) {
        b0 = std::min(n_bytes, quads * t / nt * 4); b1 = std::min(n_bytes, quads * (t + 1) / nt * 4);
    };
    auto run = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(fn, t);
        fn(0u);
        for (auto& x : th) x.join();
    };
    run([&](unsigned t) {
        uint64_t b0, b1, c = 0; slice(t, b0, b1);
        for (uint64_t p = b0; p < b1; ++p) c += lut[seq[p]] == 0xFF;
        cnt[t + 1] = c;
    });
    for (unsigned t = 0; t < nt; ++t) cnt[t + 1] += cnt[t];
    run([&](unsigned t) {
        uint64_t b0, b1, e = cnt[t]; slice(t, b0, b1);
        for (uint64_t p = b0; p < b1; p += 4) {
            uint32_t byte = 0;
            const uint64_t lim = std::min<uint64_t>(4, b1 - p);
            for (uint64_t k = 0; k < lim; ++k) {
                uint32_t c = lut[seq[p + k]];
                if (c == 0xFF) { if (e < exc_cap) { exc_pos[e] = p + k; exc_byte[e] = seq[p + k]; } ++e; c = 0; }
                byte |= c << (2 * k);
            }
            if (codes) codes[p / 4] = (uint8_t)byte;
        }
    });
    return (int64_t)cnt[nt];
}

extern "C" int b2a_seq2_unpack(const b2a_seq2* s, uint64_t first, uint64_t count, uint8_t* out)
{
    if (!s || (count && !out) || first > s->n_bytes || count > s->n_bytes - first) return B2A_ERR_ARG;
    if (count && !s->codes) return B2A_ERR_ARG;
    if (s->n_exc && (!s->exc_pos || !s->exc_byte)) return B2A_ERR_ARG;
    for (uint64_t k = 0; k < count; ++k) {
        const uint64_t p = first + k;
        out[k] = s->alphabet[(s->codes[p / 4] >> (2 * (p % 4))) & 3u];
    }
    const uint64_t* lo = std::lower_bound(s->exc_pos, s->exc_pos + s->n_exc, first);
    for (; lo != s->exc_pos + s->n_exc && *lo < first + count; ++lo) out[*lo - first] = s->exc_byte[lo - s->exc_pos];
    return B2A_OK;
}
