"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b2align.h declares, fails loudly without a GPU, and the host-only formatting/selection
entry points agree with the oracle.  No compute calls here (no GPU in the build container)."""
import ctypes as C
import json
import os
import random
import re
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
from __graft_entry__ import load_package

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = load_package()
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "hw2_kat.json")))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2align.h")).read()
    declared = sorted(set(re.findall(r"\b(b2a_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = pkg.load_library()
    for name in declared:
        assert hasattr(lib, name), f"libb2align.so does not export {name}"
    assert sorted(pkg.EXPORTS) == declared


def test_result_record_layout():
    assert pkg.RESULT_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_gpu():
    lib = pkg.load_library()
    if lib.b2a_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B2AError):
        pkg.Engine(0)


def test_render_matches_oracle_and_golden():
    for c in KAT["single"]:
        mode = ob.GLOBAL if c["mode"] == "g" else ob.LOCAL
        p, t = c["p"].encode("latin-1"), c["t"].encode("latin-1")
        a = ob.align(mode, p, t, *c["s"])
        assert pkg.render_cigar(a.ops) == c["cigar"]
        assert pkg.render_mdz(a.ops, p, t, a.start_i, a.start_j) == c["mdz"]


def test_render_rejects_small_buffer():
    lib = pkg.load_library()
    buf = C.create_string_buffer(2)
    assert lib.b2a_render_cigar(b"MMMMDDDD", 8, buf, 2) < 0


def test_select_best_first_strict_max():
    rng = random.Random(5)
    for _ in range(200):
        n = rng.randint(0, 12)
        res = np.zeros(n, dtype=pkg.RESULT_DTYPE)
        res["score"] = [rng.randint(-3, 4) for _ in range(n)]
        res["overlap"] = [rng.randint(0, 3) for _ in range(n)]
        for mode, key in ((pkg.GLOBAL, "overlap"), (pkg.LOCAL, "score")):
            want, best = -1, -1000000
            for k in range(n):
                if res[key][k] > best:
                    best, want = int(res[key][k]), k
            assert pkg.select_best(mode, res) == want


def test_aligned_strings_roundtrip():
    a = ob.align(ob.GLOBAL, b"ATCAAGCGTCGGCATATGGC", b"ATCAGCGATCATCGGCATAT", 1, -1, -1)
    ap, ar = pkg.aligned_strings(a.ops, b"ATCAAGCGTCGGCATATGGC", b"ATCAGCGATCATCGGCATAT", 0, 0)
    assert ap.replace(b"-", b"") == b"ATCAAGCGTCGGCATATGGC" and ar.replace(b"-", b"") == b"ATCAGCGATCATCGGCATAT"
    assert len(ap) == len(ar) == len(a.ops)


# ---- hw2 drop-in CLI: everything that happens before the GPU is touched ----
def run_cli(args, cwd):
    return subprocess.run([pkg.HW2_BIN] + args, cwd=cwd, capture_output=True)


def test_cli_usage_and_error_paths_match_reference(tmp_path):
    if not os.path.exists(pkg.HW2_BIN):
        pytest.skip("bin/hw2 not built")
    cases = [
        ["-g"],                                                                  # argc < 9 -> usage
        ["-g", "-p", "nope.fa", "-t", "nope2.fa", "-o", "o.txt", "-s", "1", "-1", "-1"],     # unopenable pattern file
        ["-l", "-p", "p.fa", "-t", "t2.fa", "-o", "o.txt", "-s", "1", "-1", "-1"],           # count mismatch
        ["-l", "-p", "e.fa", "-t", "e.fa", "-o", "o.txt", "-s", "1", "-1", "-1"],            # zero pairs -> empty file
        ["-g", "-p", "e.fa", "-t", "e.fa", "-o", "nodir/o.txt", "-s", "1", "-1", "-1"],      # unopenable output
        ["-p", "e.fa", "-t", "e.fa", "-o", "o2.txt", "-s", "1", "-1", "-1", "extra"],        # neither -g nor -l
    ]
    (tmp_path / "p.fa").write_bytes(b">a\nACGT\n>b\nAC\n")
    (tmp_path / "t2.fa").write_bytes(b">a\nACGT\n")
    (tmp_path / "e.fa").write_bytes(b">only header\n\n")
    for args in cases:
        for f in ("o.txt", "o2.txt"):
            if (tmp_path / f).exists():
                (tmp_path / f).unlink()
        mine = run_cli(args, tmp_path)
        outs = {f: (tmp_path / f).read_bytes() if (tmp_path / f).exists() else None for f in ("o.txt", "o2.txt")}
        if ob.have_ref():
            for f in ("o.txt", "o2.txt"):
                if (tmp_path / f).exists():
                    (tmp_path / f).unlink()
            ref = subprocess.run([ob.REF_HW2] + args, cwd=tmp_path, capture_output=True)
            refouts = {f: (tmp_path / f).read_bytes() if (tmp_path / f).exists() else None for f in ("o.txt", "o2.txt")}
            assert mine.returncode == ref.returncode, args
            assert mine.stderr.replace(pkg.HW2_BIN.encode(), b"X") == ref.stderr.replace(ob.REF_HW2.encode(), b"X"), args
            assert outs == refouts, args
        else:
            assert mine.returncode in (0, 1)


def test_upgma_newick_reproduces_trees_written_by_the_reference():
    """b2a_upgma_newick (host only) on oracle distances == the tree the unmodified hw4 binary wrote (hw4.cpp:154-237)."""
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "hw4_kat.json")))
    for c in kat["trees"] + kat["pairs"]:
        seqs = [s.encode() for s in c["seqs"]]
        dist = [ob.hw4_nw(seqs[i], seqs[j], *c["s"])[1] for i in range(len(seqs)) for j in range(i + 1, len(seqs))]
        assert pkg.upgma_newick(dist, ["s%d" % i for i in range(len(seqs))]) + "\n" == c["tree"], c
    assert pkg.upgma_newick([], ["only"]) == "only:0.0;"
    lib = pkg.load_library()
    buf = C.create_string_buffer(4)
    names = (C.c_char_p * 2)(b"a", b"b")
    d = np.array([3], dtype=np.int32)
    assert lib.b2a_upgma_newick(d.ctypes.data, 2, names, buf, 4) < 0          # buffer too small


def test_hw4_cli_error_paths_match_reference(tmp_path):
    """usage / unknown option / missing input: same stderr text and exit code as the reference binary."""
    ref = ob.REF_HW4
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/hw4 not built")
    (tmp_path / "in.fa").write_text(">a\nACGT\n")
    cases = [[], ["-i", "x"], ["-i", str(tmp_path / "nope.fa"), "-t", "t.txt", "-s", "1", "-1", "-1"],
             ["-q", "1", "-t", "t.txt", "-s", "1", "-1", "-1"],
             ["-i", str(tmp_path / "in.fa"), "-t", str(tmp_path / "nodir" / "t.txt"), "-s", "1", "-1", "-1"]]
    for args in cases:
        outs = []
        for binary in (ref, pkg.HW4_BIN):
            p = subprocess.run([binary] + args, capture_output=True, text=True, cwd=tmp_path)
            outs.append((p.returncode, p.stderr.replace(binary, "hw4")))
        if args and args[-1] == "-1" and "nodir" in args[4]:
            # reaching the writer needs the distance stage; a single sequence has no pairs, so no GPU is involved
            pass
        assert outs[0] == outs[1], (args, outs)
    # one sequence: no pairs -> runs without a GPU, byte-identical tree
    for binary, name in ((ref, "r.txt"), (pkg.HW4_BIN, "m.txt")):
        subprocess.check_call([binary, "-i", str(tmp_path / "in.fa"), "-t", str(tmp_path / name), "-s", "1", "-1", "-1"])
    assert (tmp_path / "r.txt").read_bytes() == (tmp_path / "m.txt").read_bytes() == b"a:0.0;\n"


def test_center_star_phylip_reproduces_files_written_by_the_reference():
    """b2a_center_star_phylip (host only) on oracle op lists == the file the unmodified hw3 binary wrote (hw3.cpp:231-357)."""
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "hw3_kat.json")))
    for c in kat["stars"] + kat["pairs"]:
        seqs = [s.encode() for s in c["seqs"]]
        k = len(seqs)
        sums = [0] * k
        for i in range(k):
            for j in range(i + 1, k):
                v = ob.affine_score(seqs[i], seqs[j], *c["s"])
                sums[i] += v; sums[j] += v
        centre = max(range(k), key=lambda i: (sums[i], -i))
        ops = [None if i == centre else ob.affine_align(seqs[centre], seqs[i], *c["s"])[1] for i in range(k)]
        assert pkg.center_star_phylip(["s%d" % i for i in range(k)], seqs, centre, ops) == c["phy"], c


def test_hw3_cli_messages_match_reference(tmp_path):
    """usage / unknown argument / bad score string / missing input / no or one sequence: same stdout, exit code and files."""
    ref = ob.REF_HW3
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/hw3 not built")
    (tmp_path / "one.fa").write_text("junk before\n>a b\nAC GT\nTT\n")
    (tmp_path / "none.fa").write_text("ACGT\n\n")
    cases = [[], ["-i", "x"], ["-i", "one.fa", "-q", "o.phy", "-s", "1:2:3:4"], ["-i", "one.fa", "-o", "o.phy", "-s", "1:2:3"],
             ["-i", "nope.fa", "-o", "o.phy", "-s", "5:-4:-16:-4"], ["-i", "none.fa", "-o", "o.phy", "-s", "5:-4:-16:-4"],
             ["-i", "one.fa", "-o", "o.phy", "-s", "5:-4:-16:-4"]]
    for args in cases:
        outs = []
        for binary in (ref, pkg.HW3_BIN):
            if (tmp_path / "o.phy").exists():
                (tmp_path / "o.phy").unlink()
            p = subprocess.run([binary] + args, capture_output=True, text=True, cwd=tmp_path)
            body = (tmp_path / "o.phy").read_bytes() if (tmp_path / "o.phy").exists() else None
            outs.append((p.returncode, p.stdout.replace(binary, "hw3"), p.stderr, body))
        assert outs[0] == outs[1], (args, outs)


def test_seq2_pack_unpack_round_trip(monkeypatch):
    """b2a_seq2 (the 2-bit wire format of b2a_align_batch_multi_seq2) is lossless: codes + exception list give back every byte, whatever the
    thread count of the packer, the buffer length modulo 4 and the share of bytes outside the alphabet."""
    rng = np.random.default_rng(12)
    sym = np.frombuffer(b"ACGTNacgt-\x00\xff", np.uint8)
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("B2A_HOST_THREADS", threads)
        for n in (0, 1, 2, 3, 4, 5, 63, 64, 1001, 262144 * 3 + 1):
            for p_exc in (0.0, 0.01, 1.0):
                w = np.array([1 - p_exc] * 4 + [p_exc] * 8) / (4 * (1 - p_exc) + 8 * p_exc)
                d = sym[rng.choice(12, n, p=w)]
                ps = pkg.PackedSeq(d)
                assert np.array_equal(ps.unpack(), d)
                inside = np.isin(d, sym[:4])
                assert ps.n_exc == int((~inside).sum())
                pos = ps.exc_pos[:ps.n_exc].astype(np.int64)
                assert np.array_equal(pos, np.flatnonzero(~inside)) and np.array_equal(ps.exc_byte[:ps.n_exc], d[pos])
                if n > 40:
                    assert np.array_equal(ps.unpack(n // 3, 37), d[n // 3: n // 3 + 37])
    # another alphabet order, and a repeated alphabet byte (takes its lowest code)
    d = np.frombuffer(b"TTGACCAGTNAC", np.uint8)
    assert np.array_equal(pkg.PackedSeq(d, alphabet=b"TGCA").unpack(), d)
    assert np.array_equal(pkg.PackedSeq(d, alphabet=b"AACG").unpack(), d)
    lib = pkg.load_library()
    ps = pkg.PackedSeq(d)
    assert lib.b2a_seq2_unpack(C.byref(ps.c), 5, 100, np.zeros(200, np.uint8).ctypes.data) == -1          # beyond the buffer
    assert lib.b2a_seq2_pack(d.ctypes.data, d.size, None, None, None, None, 0) == -1


def test_find_anchors_chain_properties():
    """b2a_find_anchors (host only): every anchor is an exact match, the chain ascends strictly in both sequences without
    overlaps, respects the spacing, and degenerate inputs give an empty chain instead of an error."""
    rng = random.Random(9)
    base = bytes(rng.choice(b"ACGT") for _ in range(30000))
    mut = bytearray()
    for ch in base:
        r = rng.random()
        if r < 0.01:
            continue
        if r < 0.02:
            mut.append(rng.choice(b"ACGT"))
        mut.append(rng.choice(b"ACGT") if rng.random() < 0.06 else ch)
    mut = bytes(mut)
    for k, spacing in ((12, 64), (16, 256), (20, 1000), (16, 1)):
        a = pkg.find_anchors(mut, base, k, spacing)
        assert len(a) > 30000 // max(spacing, 3 * k) // 4
        for x in a:
            assert mut[x["i"]:x["i"] + x["len"]] == base[x["j"]:x["j"] + x["len"]] and x["len"] == k
        i, j = a["i"].astype(np.int64), a["j"].astype(np.int64)
        assert np.all(np.diff(i) >= max(spacing, k)) and np.all(np.diff(j) >= k)
    assert len(pkg.find_anchors(b"ACGT", base, 16, 64)) == 0                       # shorter than k
    assert len(pkg.find_anchors(b"", b"", 16, 64)) == 0
    assert len(pkg.find_anchors(b"ACGT" * 500, b"ACGT" * 400, 16, 64)) == 0        # no unique k-mer
    assert len(pkg.find_anchors(base[:5000], bytes(rng.choice(b"ACGT") for _ in range(5000)), 16, 64)) == 0    # unrelated sequences
    # a reversed block cannot be chained with its surroundings: the chain stays monotone
    t2 = base[:10000] + base[20000:30000] + base[10000:20000]
    a = pkg.find_anchors(base, t2, 16, 128)
    assert np.all(np.diff(a["j"].astype(np.int64)) > 0) and len(a) > 50
    lib = pkg.load_library()
    assert lib.b2a_find_anchors(None, 10, None, 10, 16, 1, None, 0) == -1
    assert lib.b2a_find_anchors(base, 100, base, 100, 0, 1, None, 0) == -1
    assert lib.b2a_align_anchored(None, None, None, 0, None, 0, None, 0, None, None, 0) == -1


def test_seq2_and_anchor_properties_hypothesis():
    """Property tests of the two host-only additions: pack -> unpack is the identity for arbitrary bytes and alphabets; every anchor chain
    consists of exact matches ascending in both coordinates; forcing a chain through identical sequences covers them on the diagonal."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.binary(min_size=0, max_size=600), st.binary(min_size=4, max_size=4))
    def roundtrip(data, alphabet):
        d = np.frombuffer(data, np.uint8)
        ps = pkg.PackedSeq(d, alphabet=alphabet)
        assert ps.unpack().tobytes() == data
        assert ps.n_exc == sum(1 for b in data if b not in alphabet)

    @settings(max_examples=60, deadline=None)
    @given(st.text(alphabet="ACGT", min_size=0, max_size=400), st.text(alphabet="ACGT", min_size=0, max_size=400),
           st.integers(min_value=4, max_value=12), st.integers(min_value=1, max_value=64))
    def chain(p, t, k, spacing):
        p, t = p.encode(), t.encode()
        a = pkg.find_anchors(p, t, k, spacing)
        pi = tj = 0
        for x in a:
            i, j, ln = int(x["i"]), int(x["j"]), int(x["len"])
            assert ln == k and p[i:i + ln] == t[j:j + ln] and i >= pi and j >= tj
            pi, tj = i + ln, j + ln

    roundtrip()
    chain()
    rng = random.Random(2)
    s_ = bytes(rng.choice(b"ACGT") for _ in range(5000))
    a = pkg.find_anchors(s_, s_, 16, 100)
    assert len(a) >= 40 and np.array_equal(a["i"], a["j"])


def test_header_is_plain_c_and_links(tmp_path):
    """include/b2align.h compiles as strict C99 (the boundary is a C ABI: no C++ types, no default arguments) and a C program linked against
    libb2align.so can call the host-only entry points; the computing ones report the missing GPU instead of falling back."""
    src = tmp_path / "abi.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "b2align.h"
int main(void) {
    const uint8_t acgt[4] = {'A', 'C', 'G', 'T'};
    const uint8_t seq[] = "ACGTNACGTACGTTTGA";
    uint8_t codes[8], exc_byte[4], back[17];
    uint64_t exc_pos[4];
    int64_t ne = b2a_seq2_pack(seq, 17, acgt, codes, exc_pos, exc_byte, 4);
    b2a_seq2 s;
    char cigar[64];
    b2a_anchor a[4];
    b2a_result r;
    memset(&s, 0, sizeof s);
    s.codes = codes; s.n_bytes = 17; memcpy(s.alphabet, acgt, 4); s.exc_pos = exc_pos; s.exc_byte = exc_byte; s.n_exc = (uint64_t)ne;
    if (ne != 1 || exc_pos[0] != 4 || exc_byte[0] != 'N') return 1;
    if (b2a_seq2_unpack(&s, 0, 17, back) != B2A_OK || memcmp(back, seq, 17) != 0) return 2;
    if (b2a_render_cigar("MMIDM", 5, cigar, sizeof cigar) != 8 || strcmp(cigar, "1M1D1I2M") != 0) return 3;
    if (b2a_find_anchors(seq, 17, seq, 17, 4, 1, a, 4) < 0) return 4;
    if (sizeof(b2a_result) != 32 || sizeof(b2a_anchor) != 12 || sizeof r != 32) return 5;
    if (b2a_device_count() == 0 && b2a_create(0) != NULL) return 6;          /* no GPU: no context, no fallback */
    printf("ok %d\n", B2A_VERSION);
    return 0;
}
""")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lb2align", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), (out.returncode, out.stdout, out.stderr)


def test_anchor_chains_keep_the_optimal_score_oracle_only():
    """The f4 semantics on the CPU, no engine involved: the chain b2a_find_anchors returns for a mutated copy (2-25 % substitutions, 1 % indels
    each way) never costs score -- hw2's NW restated by the oracle on every stretch between anchors, plus the anchors' matches, equals the
    oracle's unconstrained NW score.  (200 pairs of 2-5 kb: 200 / 200 equal, 176 / 200 with the identical op list; 30 of them run here.)"""
    rng = random.Random(2026)

    def mutate(seq, psub):
        out = bytearray()
        for ch in seq:
            r = rng.random()
            if r < 0.01:
                continue
            if r < 0.02:
                out.append(rng.choice(b"ACGT"))
            out.append(rng.choice(b"ACGT") if rng.random() < psub else ch)
        return bytes(out)

    for psub in (0.02, 0.05, 0.10, 0.15, 0.25):
        for _ in range(6):
            t = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1500, 3000)))
            p = mutate(t, psub)
            anchors = pkg.find_anchors(p, t, 16, 256)
            score, pi, tj = 0, 0, 0
            for x in list(anchors) + [None]:
                pe, te = (int(x["i"]), int(x["j"])) if x is not None else (len(p), len(t))
                score += ob.align(ob.GLOBAL, p[pi:pe], t[tj:te], 1, -1, -1).score
                if x is not None:
                    score += int(x["len"])
                    pi, tj = pe + int(x["len"]), te + int(x["len"])
            full = ob.align(ob.GLOBAL, p, t, 1, -1, -1).score
            assert score <= full
            assert score == full, (psub, len(anchors), score, full)
