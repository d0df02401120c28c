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
    std::memset(lut, 0x80, sizeof lut);                                  // 0x80 = not in the alphabet
    for (int c = 3; c >= 0; --c) lut[alphabet[c]] = (uint8_t)c;          // a repeated alphabet byte takes its lowest code
    // ONE pass: slices of whole code bytes, one per thread; a thread writes its codes in place and collects its exceptions in a list of
    // its own (four bytes at a time, the exception test once per four), the lists are concatenated afterwards.
    const uint64_t quads = (n_bytes + 3) / 4;
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (const char* e = std::getenv("B2A_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(1, quads / 65536));
    struct Exc { uint64_t pos; uint8_t byte; };
    std::vector<std::vector<Exc>> found(nt);
    auto work = [&](unsigned t) {
        const uint64_t b0 = std::min(n_bytes, quads * t / nt * 4), b1 = std::min(n_bytes, quads * (t + 1) / nt * 4);
        std::vector<Exc>& ex = found[t];
        uint64_t p = b0;
        for (; p + 4 <= b1; p += 4) {
            uint32_t l0 = lut[seq[p]], l1 = lut[seq[p + 1]], l2 = lut[seq[p + 2]], l3 = lut[seq[p + 3]];
            if ((l0 | l1 | l2 | l3) & 0x80u) {
                if (l0 & 0x80u) { ex.push_back(Exc{p, seq[p]}); l0 = 0; }
                if (l1 & 0x80u) { ex.push_back(Exc{p + 1, seq[p + 1]}); l1 = 0; }
                if (l2 & 0x80u) { ex.push_back(Exc{p + 2, seq[p + 2]}); l2 = 0; }
                if (l3 & 0x80u) { ex.push_back(Exc{p + 3, seq[p + 3]}); l3 = 0; }
            }
            if (codes) codes[p / 4] = (uint8_t)(l0 | (l1 << 2) | (l2 << 4) | (l3 << 6));
        }
        if (p < b1) {                                                    // the buffer's last, partial code byte
            uint32_t byte = 0;
            for (uint64_t k = 0; p + k < b1; ++k) {
                uint32_t c = lut[seq[p + k]];
                if (c & 0x80u) { ex.push_back(Exc{p + k, seq[p + k]}); c = 0; }
                byte |= c << (2 * k);
            }
            if (codes) codes[p / 4] = (uint8_t)byte;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0u);
        for (auto& x : th) x.join();
    }
    uint64_t total = 0;
    for (unsigned t = 0; t < nt; ++t)
        for (const Exc& e : found[t]) { if (total < exc_cap) { exc_pos[total] = e.pos; exc_byte[total] = e.byte; } ++total; }
    return (int64_t)total;
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

// ---- seed-anchored global alignment (not in the reference's code, see include/b2align.h) ---------------------------
// Built on the public batch call only: the stretches between anchors are ordinary pairs of b2a_align_batch.

extern "C" int64_t b2a_find_anchors(const uint8_t* pattern, uint64_t m, const uint8_t* text, uint64_t n, uint32_t k, uint32_t spacing,
                                    b2a_anchor* out, uint64_t cap)
{
    if ((!pattern && m) || (!text && n) || k == 0 || (cap && !out) || m > 0xFFFFFFF0ull || n > 0xFFFFFFF0ull) return B2A_ERR_ARG;
    if (m < k || n < k) return 0;
    // One open-addressing table per side: k-mer hash -> (first position, occurrences).  A k-mer is a candidate iff it occurs once on
    // each side and the bytes agree.  Walking the pattern in order yields the candidates sorted by i (no sort of 1e5 records on the path).
    // The table accesses are cache misses (4 MB per 100 kb side), so only a hash-selected 1/2 .. 1/16 of the k-mers is considered when
    // the spacing leaves room for it (>= 32 candidates per kept anchor either way); a k-mer is selected on both sides or on neither.
    struct Slot { uint64_t h; uint32_t pos, cnt; };
    uint64_t smask = 0;
    while (smask < 15 && (smask + 1) * 2 * 32 <= (uint64_t)spacing) smask = smask * 2 + 1;
    const uint64_t B = 0x9E3779B97F4A7C15ull;
    uint64_t top = 1;
    for (uint32_t t = 1; t < k; ++t) top *= B;
    auto roll = [&](const uint8_t* s, uint64_t len, auto&& visit) {           // visit(hash, position) for every k-mer, in order
        uint64_t h = 0;
        for (uint32_t t = 0; t < k; ++t) h = h * B + (s[t] + 1u);
        visit(h, (uint32_t)0);
        for (uint64_t p = 1; p + k <= len; ++p) {
            h = (h - (s[p - 1] + 1u) * top) * B + (s[p + k - 1] + 1u);
            visit(h, (uint32_t)p);
        }
    };
    struct Table {
        std::vector<Slot> slots; int shift = 0;
        void init(uint64_t n_keys) { int bits = 4; while ((1ull << bits) < 2 * n_keys) ++bits; shift = 64 - bits; slots.assign(1ull << bits, Slot{0, 0, 0}); }
        Slot& find(uint64_t h) {
            const uint64_t mask = slots.size() - 1;
            for (uint64_t x = (h * 0xD6E8FEB86659FD93ull) >> shift;; x = (x + 1) & mask) { Slot& sl = slots[x]; if (sl.cnt == 0 || sl.h == h) return sl; }
        }
    };
    Table tp_, tt_;
    auto build = [&](Table& tb, const uint8_t* s, uint64_t len) {
        uint64_t selected = 0;                                               // exact, so the table can never fill up
        roll(s, len, [&](uint64_t h, uint32_t) { selected += ((h >> 24) & smask) == 0; });
        tb.init(selected + 1);
        roll(s, len, [&](uint64_t h, uint32_t pos) { if ((h >> 24) & smask) return; Slot& sl = tb.find(h); if (sl.cnt++ == 0) { sl.h = h; sl.pos = pos; } });
    };
    {
        std::thread th([&]() { build(tp_, pattern, m); });
        build(tt_, text, n);
        th.join();
    }
    std::vector<b2a_anchor> cand;
    roll(pattern, m, [&](uint64_t h, uint32_t pos) {
        if (((h >> 24) & smask) || tp_.find(h).cnt != 1) return;
        const Slot& st = tt_.find(h);
        if (st.cnt == 1 && std::memcmp(pattern + pos, text + st.pos, k) == 0) cand.push_back(b2a_anchor{pos, st.pos, k});
    });
    // longest chain with strictly increasing j (patience sorting with predecessor links); i ascends strictly already
    std::vector<uint32_t> tail, prev(cand.size(), 0xFFFFFFFFu);
    for (uint32_t c = 0; c < cand.size(); ++c) {
        auto it = std::lower_bound(tail.begin(), tail.end(), cand[c].j, [&](uint32_t t, uint32_t j) { return cand[t].j < j; });
        if (it != tail.begin()) prev[c] = *(it - 1);
        if (it == tail.end()) tail.push_back(c); else *it = c;
    }
    std::vector<b2a_anchor> chain;
    for (uint32_t c = tail.empty() ? 0xFFFFFFFFu : tail.back(); c != 0xFFFFFFFFu; c = prev[c]) chain.push_back(cand[c]);
    std::reverse(chain.begin(), chain.end());
    // thin: consecutive anchors start >= max(spacing, k) pattern bases apart and do not overlap in the text either
    uint64_t kept = 0;
    b2a_anchor last{0, 0, 0};
    for (const b2a_anchor& a : chain) {
        if (kept && (a.i < last.i + std::max(spacing, k) || a.j < last.j + last.len)) continue;
        if (kept < cap) out[kept] = a;
        last = a; ++kept;
    }
    return (int64_t)kept;
}

extern "C" int64_t b2a_align_anchored(b2a_ctx* ctx, const b2a_params* prm, const uint8_t* pattern, uint64_t m, const uint8_t* text, uint64_t n,
                                      const b2a_anchor* anchors, uint64_t n_anchors, b2a_result* result, char* ops, uint64_t ops_cap)
{
    if (!ctx || !prm || !result || (!pattern && m) || (!text && n) || (n_anchors && !anchors)) return B2A_ERR_ARG;
    if (prm->mode != B2A_MODE_GLOBAL || m + n >= 0x7FFFFFF0ull) return B2A_ERR_ARG;
    // stretch s = what lies between anchor s-1 and anchor s (n_anchors + 1 stretches, some possibly empty)
    const uint64_t ns = n_anchors + 1;
    uint64_t pi = 0, tj = 0, anchored = 0;
    for (uint64_t s = 0; s < n_anchors; ++s) {
        const b2a_anchor& a = anchors[s];
        if (a.i < pi || a.j < tj || (uint64_t)a.i + a.len > m || (uint64_t)a.j + a.len > n || a.len == 0 ||
            std::memcmp(pattern + a.i, text + a.j, a.len) != 0) return B2A_ERR_ARG;
        pi = (uint64_t)a.i + a.len; tj = (uint64_t)a.j + a.len; anchored += a.len;
    }
    // pair s of the batch = stretch s: the stretches copied back to back (the anchors' bytes are left out: an exact match needs no DP)
    std::vector<uint8_t> pb, tb;
    pb.reserve(m - anchored); tb.reserve(n - anchored);
    std::vector<uint64_t> spo(ns + 1, 0), sto(ns + 1, 0);
    pi = 0; tj = 0;
    for (uint64_t s = 0; s < ns; ++s) {
        const uint64_t pe = s < n_anchors ? anchors[s].i : m, te = s < n_anchors ? anchors[s].j : n;
        pb.insert(pb.end(), pattern + pi, pattern + pe); tb.insert(tb.end(), text + tj, text + te);
        spo[s + 1] = pb.size(); sto[s + 1] = tb.size();
        if (s < n_anchors) { pi = pe + anchors[s].len; tj = te + anchors[s].len; }
    }
    b2a_params q = *prm;
    q.flags = (q.flags & B2A_TIE_HW4) | B2A_WANT_OPS;
    b2a_set_ops_sink(ctx, nullptr, 0, 0);                     // the stretches' op lists are this call's own business, not the caller's sink's
    std::vector<b2a_result> res(ns);
    int rc = b2a_align_batch(ctx, &q, pb.data(), spo.data(), tb.data(), sto.data(), ns, res.data());
    if (rc != B2A_OK) return rc;
    const int64_t nw = b2a_copy_ops(ctx, nullptr, 0, nullptr);
    if (nw < 0) return nw;
    std::vector<uint32_t> words((size_t)nw + 1);
    std::vector<uint64_t> woff(ns + 1);
    const int64_t got = b2a_copy_ops(ctx, words.data(), (uint64_t)nw, woff.data());
    if (got < 0) return got;
    // assemble in TRACEBACK order (alignment end -> start): last stretch first, each stretch's own list as it is, anchors as runs of 'M'
    b2a_result r{};
    int64_t score = (int64_t)prm->match * (int64_t)anchored;
    uint64_t nops = 0;
    static const char L[4] = {'M', 'D', 'I', '?'};
    auto put = [&](char c) { if (ops && nops < ops_cap) ops[nops] = c; ++nops; };
    for (uint64_t s = ns; s-- > 0;) {
        const uint32_t* w = words.data() + woff[s];
        for (uint32_t t = 0; t < res[s].n_ops; ++t) put(L[(w[t >> 4] >> (2 * (t & 15))) & 3u]);
        score += res[s].score;
        if (s > 0) for (uint32_t t = 0; t < anchors[s - 1].len; ++t) put('M');
    }
    // overlapLongestExactMatch over the whole alignment (hw2.cpp:267-278), from the op list in alignment order; hw4's distance with B2A_TIE_HW4
    int best = 0, cur = 0; uint64_t i = 0, j = 0, mism = 0, gaps = 0;
    if (ops && nops <= ops_cap) {
        for (uint64_t t = nops; t-- > 0;) {
            const char c = ops[t];
            if (c == 'M') { const bool eq = pattern[i] == text[j]; cur = (eq && pattern[i] != '-') ? cur + 1 : 0; mism += !eq; ++i; ++j; }
            else { cur = 0; ++gaps; if (c == 'D') ++i; else ++j; }
            best = std::max(best, cur);
        }
        r.overlap = (prm->flags & B2A_TIE_HW4) ? (int32_t)(mism + gaps) : best;
    } else r.overlap = -1;                                   // not computed without the op list
    r.score = (int32_t)score; r.end_i = (uint32_t)m; r.end_j = (uint32_t)n; r.start_i = 0; r.start_j = 0;
    r.n_ops = (uint32_t)nops; r.path = 3;
    *result = r;
    return (int64_t)nops;
}
