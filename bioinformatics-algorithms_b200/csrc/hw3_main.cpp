// hw3_main.cpp -- drop-in replacement for the reference's centre-star MSA program
// (Multiple_Sequence_Alignment/hw3.cpp:169-368).
//
//   hw3 -i input.fasta -o output.phy -s match:mismatch:gapOpen:gapExtend
//
// Same argv grammar, FASTA rules (hw3.cpp:137-167), messages (all on stdout, exit code 0 except an unopenable
// input), special cases (no / one sequence) and output bytes.  Both alignment stages run on the GPUs through the C
// ABI, sharded over every visible device: the all-vs-all score-only distance stage (hw3.cpp:231-241 ->
// b2a_affine_star_scores) and the centre-vs-others alignments with traceback (hw3.cpp:259-266 ->
// b2a_affine_align_batch); the merge + PHYLIP text (hw3.cpp:253-357) is host code behind b2a_center_star_phylip.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "b2align.h"

int main(int argc, char** argv)
{
    if (argc < 7) {
        std::cout << "Usage: " << argv[0] << " -i input.fasta -o output.phy -s matchScore:mismatchScore:gapOpeningScore:gapExtensionScore" << std::endl;
        return 0;
    }
    std::string in_path, out_path, scores;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-i" && i + 1 < argc) in_path = argv[++i];
        else if (a == "-o" && i + 1 < argc) out_path = argv[++i];
        else if (a == "-s" && i + 1 < argc) scores = argv[++i];
        else { std::cout << "Unknown argument: " << a << std::endl; return 0; }
    }
    std::vector<int> val;
    {
        std::stringstream ss(scores);
        std::string tok;
        while (std::getline(ss, tok, ':')) val.push_back(std::stoi(tok));
        if (val.size() != 4) { std::cout << "Error: Score must have four values separated by ':'" << std::endl; return 0; }
    }
    const int match = val[0], mismatch = val[1], gopen = val[2], gext = val[3];

    // hw3.cpp:137-167: a record is flushed when the NEXT non-empty header arrives, and only then is the sequence buffer
    // cleared -- so text in front of the first header joins the first record and a record with an empty header melts
    // into its successor; whitespace inside sequence lines is dropped, headers are kept verbatim after '>'
    std::ifstream in(in_path.c_str());
    if (!in) { std::cout << "Error: Could not open file " << in_path << std::endl; return 1; }
    std::vector<std::string> ids, seqs;
    {
        std::string line, header, seq;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            if (line[0] == '>') {
                if (!header.empty()) { ids.push_back(header); seqs.push_back(seq); seq.clear(); }
                header = line.substr(1);
            } else for (char c : line) if (!std::isspace((unsigned char)c)) seq.push_back(c);
        }
        if (!header.empty()) { ids.push_back(header); seqs.push_back(seq); }
    }
    const uint32_t n = (uint32_t)ids.size();
    if (n == 0) { std::cout << "No sequences found in " << in_path << std::endl; return 0; }
    if (n == 1) {                                                               // hw3.cpp:221-228
        std::ofstream out(out_path.c_str());
        out << "1 " << seqs[0].size() << "\n" << ids[0] << " " << seqs[0] << "\n";
        return 0;
    }

    // a CUDA context costs ~1 s, driver initialisation for 8 visible GPUs several seconds; one GPU does ~2.5e12 affine cells
    // per second: spread out only from ~1e12 cells per device on, and show a one-device job one device
    double cells = 0;
    for (uint32_t i = 0; i < n; ++i) for (uint32_t j = i + 1; j < n; ++j) cells += (double)seqs[i].size() * (double)seqs[j].size();
    const bool all_gpus = std::getenv("B2A_ALL_GPUS") != nullptr;
    if (!all_gpus && cells < 2e12) setenv("CUDA_VISIBLE_DEVICES", "0", 0);           // no-op if the user set it
    int ndev = b2a_device_count();
    if (ndev <= 0) { std::cout << "Error: no usable CUDA device (this build has no CPU alignment path)" << std::endl; return 1; }
    if (!all_gpus) ndev = (int)std::max(1.0, std::min((double)ndev, cells / 1e12));
    std::vector<uint8_t> all;
    std::vector<uint64_t> off{0};
    for (const std::string& s : seqs) { all.insert(all.end(), s.begin(), s.end()); off.push_back(all.size()); }

    // ---- stage 1: star scores (hw3.cpp:231-241), pair ranges over the devices, partial sums added ----
    const uint32_t total = n * (n - 1) / 2;
    const int nd1 = (int)std::min<uint32_t>((uint32_t)ndev, total);
    std::vector<std::vector<int32_t>> part(nd1, std::vector<int32_t>(n, 0));
    std::vector<int> rc(ndev, 0);
    std::vector<std::string> err(ndev);
    std::vector<b2a_ctx*> ctxs(ndev, nullptr);
    {
        std::vector<std::thread> th;
        for (int d = 0; d < nd1; ++d)
            th.emplace_back([&, d]() {
                ctxs[d] = b2a_create(d);
                if (!ctxs[d]) { rc[d] = B2A_ERR_CUDA; err[d] = "cannot create a context on device " + std::to_string(d); return; }
                const uint32_t first = (uint32_t)((uint64_t)total * d / nd1), count = (uint32_t)((uint64_t)total * (d + 1) / nd1) - first;
                rc[d] = b2a_affine_star_scores(ctxs[d], match, mismatch, gopen, gext, all.data(), off.data(), n, first, count,
                                               nullptr, part[d].data(), nullptr);
                if (rc[d] != B2A_OK) err[d] = b2a_last_error(ctxs[d]);
            });
        for (auto& t : th) t.join();
    }
    auto bail = [&]() { for (b2a_ctx* c : ctxs) b2a_destroy(c); };
    for (int d = 0; d < nd1; ++d) if (rc[d] != B2A_OK) { std::cout << "Error: alignment engine failed: " << err[d] << std::endl; bail(); return 1; }
    std::vector<int> sums(n, 0);
    for (int d = 0; d < nd1; ++d) for (uint32_t i = 0; i < n; ++i) sums[i] += part[d][i];
    uint32_t centre = 0;                                                        // hw3.cpp:243-251: first strict maximum
    for (uint32_t i = 1; i < n; ++i) if (sums[i] > sums[centre]) centre = i;

    // ---- stage 2: centre against every other sequence, with traceback (hw3.cpp:259-266) ----
    std::vector<uint32_t> others;
    for (uint32_t i = 0; i < n; ++i) if (i != centre) others.push_back(i);
    std::vector<std::string> ops(n);
    const int nd2 = (int)std::min<size_t>((size_t)ndev, others.size());
    {
        std::vector<std::thread> th;
        for (int d = 0; d < nd2; ++d)
            th.emplace_back([&, d]() {
                if (!ctxs[d]) ctxs[d] = b2a_create(d);
                if (!ctxs[d]) { rc[d] = B2A_ERR_CUDA; err[d] = "cannot create a context on device " + std::to_string(d); return; }
                const size_t first = others.size() * d / nd2, count = others.size() * (d + 1) / nd2 - first;
                std::vector<uint8_t> pat, txt;
                std::vector<uint64_t> po{0}, to{0};
                for (size_t k = 0; k < count; ++k) {
                    const std::string& o = seqs[others[first + k]];
                    pat.insert(pat.end(), seqs[centre].begin(), seqs[centre].end()); po.push_back(pat.size());
                    txt.insert(txt.end(), o.begin(), o.end()); to.push_back(txt.size());
                }
                std::vector<int32_t> sc(count);
                std::vector<uint32_t> nops(count);
                rc[d] = b2a_affine_align_batch(ctxs[d], match, mismatch, gopen, gext, pat.data(), po.data(), txt.data(), to.data(), count, sc.data(), nops.data());
                for (size_t k = 0; k < count && rc[d] == B2A_OK; ++k) {
                    std::string& dst = ops[others[first + k]];
                    dst.resize(nops[k]);
                    if (b2a_affine_fetch_ops(ctxs[d], k, dst.empty() ? nullptr : &dst[0], nops[k]) < 0 && nops[k]) rc[d] = B2A_ERR_CUDA;
                }
                if (rc[d] != B2A_OK) err[d] = b2a_last_error(ctxs[d]);
            });
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < nd2; ++d) if (rc[d] != B2A_OK) { std::cout << "Error: alignment engine failed: " << err[d] << std::endl; bail(); return 1; }
    bail();

    // ---- merge + PHYLIP (hw3.cpp:253-357) ----
    std::vector<const char*> names(n), opp(n);
    std::vector<const uint8_t*> sp(n);
    std::vector<uint64_t> sl(n), on(n);
    uint64_t width = seqs[centre].size() + 1;
    for (uint32_t i = 0; i < n; ++i) {
        names[i] = ids[i].c_str(); sp[i] = (const uint8_t*)seqs[i].data(); sl[i] = seqs[i].size();
        opp[i] = ops[i].data(); on[i] = ops[i].size();
        width += seqs[i].size();
    }
    std::vector<char> buf((size_t)(n * (width + width / 10 + 16) + 64));
    const int64_t len = b2a_center_star_phylip(n, centre, names.data(), sp.data(), sl.data(), opp.data(), on.data(), buf.data(), buf.size());
    if (len < 0) { std::cout << "Error: alignment assembly failed" << std::endl; return 1; }
    std::ofstream out(out_path.c_str(), std::ios::binary);
    if (!out) { std::cout << "Error: Could not open output file " << out_path << std::endl; return 0; }
    out.write(buf.data(), len);
    out.close();
    return 0;
}
