// hw4_main.cpp -- drop-in replacement for the reference's UPGMA program (hw4/hw4.cpp:75-240).
//
//   hw4 -i <input.fasta> -t <tree.txt> -s <match> <mismatch> <gap>
//
// Same argv grammar, FASTA rules, messages, exit codes and output bytes.  The all-vs-all distance stage
// (hw4.cpp:137-159: n(n-1)/2 Needleman-Wunsch alignments with hw4's own tie order d > u > l, distance =
// mismatch + gap columns) runs on the GPUs through the C ABI (B2A_TIE_HW4), pair-sharded over every visible
// device; UPGMA + Newick (hw4.cpp:162-237) is host code behind b2a_upgma_newick.  No CPU alignment path.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "b2align.h"

int main(int argc, char** argv)
{
    if (argc < 7) {
        std::cerr << "Usage: " << argv[0] << " -i <input.fasta> -t <tree.txt> -s <match> <mismatch> <gap>\n";
        return 1;
    }
    std::string in_path = "input.fasta", out_path = "tree.txt";
    int match = 1, mismatch = -1, gap = -1;
    for (int i = 1; i < argc; ++i) {
        const std::string opt = argv[i];
        if (opt == "-i" && i + 1 < argc) in_path = argv[++i];
        else if (opt == "-t" && i + 1 < argc) out_path = argv[++i];
        else if (opt == "-s" && i + 3 < argc) { match = std::stoi(argv[++i]); mismatch = std::stoi(argv[++i]); gap = std::stoi(argv[++i]); }
        else { std::cerr << "Unknown option: " << opt << '\n'; return 1; }
    }
    std::ifstream in(in_path.c_str());
    if (!in) { std::cerr << "Error opening input file: " << in_path << '\n'; return 1; }

    // hw4.cpp:104-131: ids kept, records with an empty id dropped, one trailing CR stripped, blank lines skipped
    std::vector<std::string> ids, seqs;
    std::string line, id, seq;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            if (!id.empty()) { ids.push_back(id); seqs.push_back(seq); }
            id = line.substr(1);
            seq.clear();
        } else seq += line;
    }
    if (!id.empty()) { ids.push_back(id); seqs.push_back(seq); }
    in.close();
    const uint32_t n = (uint32_t)ids.size();
    if (n == 0) { std::cerr << "Error: no sequences in " << in_path << '\n'; return 1; }   // (the reference crashes here)

    // pair k = (i, j), i < j, row-major: pattern = sequence i (rows), text = sequence j (columns)
    std::vector<uint8_t> pat, txt;
    std::vector<uint64_t> po{0}, to{0};
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = i + 1; j < n; ++j) {
            pat.insert(pat.end(), seqs[i].begin(), seqs[i].end()); po.push_back(pat.size());
            txt.insert(txt.end(), seqs[j].begin(), seqs[j].end()); to.push_back(txt.size());
        }
    const uint64_t n_pairs = po.size() - 1;
    std::vector<b2a_result> results(n_pairs);
    if (n_pairs > 0) {
        // a CUDA context costs ~1 s, driver initialisation for 8 visible GPUs several seconds: spread out only from ~1e12
        // cells per device on, and show a one-device job one device
        double cells = 0;
        for (uint64_t k = 0; k < n_pairs; ++k) cells += (double)(po[k + 1] - po[k]) * (double)(to[k + 1] - to[k]);
        const bool all_gpus = std::getenv("B2A_ALL_GPUS") != nullptr;
        if (!all_gpus && cells < 2e12) setenv("CUDA_VISIBLE_DEVICES", "0", 0);       // no-op if the user set it
        int ndev = b2a_device_count();
        if (ndev <= 0) { std::cerr << "Error: no usable CUDA device (this build has no CPU alignment path)" << std::endl; return 1; }
        if ((uint64_t)ndev > n_pairs) ndev = (int)n_pairs;
        if (!all_gpus) ndev = (int)std::max(1.0, std::min((double)ndev, cells / 1e12));
        std::vector<int> rc(ndev, 0);
        std::vector<std::string> err(ndev);
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; ++d)
            th.emplace_back([&, d]() {
                const uint64_t first = n_pairs * d / ndev, count = n_pairs * (d + 1) / ndev - first;
                b2a_ctx* ctx = b2a_create(d);
                if (!ctx) { rc[d] = B2A_ERR_CUDA; err[d] = "cannot create a context on device " + std::to_string(d); return; }
                std::vector<uint64_t> p(count + 1), t(count + 1);
                for (uint64_t k = 0; k <= count; ++k) { p[k] = po[first + k] - po[first]; t[k] = to[first + k] - to[first]; }
                b2a_params prm{B2A_MODE_GLOBAL, match, mismatch, gap, B2A_TIE_HW4};
                rc[d] = b2a_align_batch(ctx, &prm, pat.data() + po[first], p.data(), txt.data() + to[first], t.data(), count, results.data() + first);
                if (rc[d] != B2A_OK) err[d] = b2a_last_error(ctx);
                b2a_destroy(ctx);
            });
        for (auto& t : th) t.join();
        for (int d = 0; d < ndev; ++d)
            if (rc[d] != B2A_OK) { std::cerr << "Error: alignment engine failed: " << err[d] << std::endl; return 1; }
    }
    std::vector<int32_t> dist(n_pairs);
    for (uint64_t k = 0; k < n_pairs; ++k) dist[k] = results[k].overlap;          // hw4.cpp:146-152 (B2A_TIE_HW4)
    std::vector<const char*> names(n);
    size_t cap = 64;
    for (uint32_t i = 0; i < n; ++i) { names[i] = ids[i].c_str(); cap += ids[i].size() + 64; }
    std::vector<char> buf(cap);
    if (b2a_upgma_newick(dist.data(), n, names.data(), buf.data(), buf.size()) < 0) { std::cerr << "Error: tree construction failed" << std::endl; return 1; }

    std::ofstream out(out_path.c_str());
    if (!out) { std::cerr << "Error opening output file: " << out_path << '\n'; return 1; }
    out << buf.data() << std::endl;
    out.close();
    return 0;
}
