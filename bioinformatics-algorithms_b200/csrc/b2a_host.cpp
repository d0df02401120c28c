// b2a_host.cpp -- host-only pieces of the C ABI: result formatting and batch selection.
// (prepareCigarString hw2.cpp:59-78, prepareMDZString hw2.cpp:80-116, selection hw2.cpp:326-357.)
// Works on the op list in traceback order plus the raw sequences; no aligned strings are built.
#include <cstdint>
#include <cstdio>
#include <string>

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
