"""Pins oracle/hw2_oracle.c (the CPU restatement) against the reference's own outputs.

  * tests/golden/hw2_kat.json -- produced by the UNMODIFIED reference binary
    (tests/golden/make_golden.py), including the shipped global.txt/local.txt;
  * live differential runs against oracle/_ref/hw2 when it is present (build
    container; it also travels to the GPU box as a prebuilt file).
"""
import json
import os
import random

import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "hw2_kat.json")))
MODE = {"g": ob.GLOBAL, "l": ob.LOCAL}


def test_golden_single_pairs():
    for c in KAT["single"]:
        a = ob.align(MODE[c["mode"]], c["p"].encode("latin-1"), c["t"].encode("latin-1"), *c["s"])
        assert (a.score, a.cigar, a.mdz) == (c["score"], c["cigar"], c["mdz"]), c


def test_golden_multi_pair_selection():
    for c in KAT["multi"]:
        out = ob.render_file(MODE[c["mode"]], [x.encode() for x in c["patterns"]], [x.encode() for x in c["texts"]], *c["s"])
        assert out.decode("latin-1") == c["output"], c


def test_shipped_fixture_byte_exact():
    sh = KAT["shipped"]
    ps, ts = [x.encode() for x in sh["patterns"]], [x.encode() for x in sh["texts"]]
    assert ob.render_file(ob.GLOBAL, ps, ts, *sh["s"]).decode() == sh["global_txt"]
    assert ob.render_file(ob.LOCAL, ps, ts, *sh["s"]).decode() == sh["local_txt"]


def test_score_only_matches_full():
    rng = random.Random(7)
    for _ in range(200):
        p = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 60)))
        t = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 80)))
        s = rng.choice([(1, -1, -1), (2, -3, -4), (5, -4, -16)])
        for mode in (ob.GLOBAL, ob.LOCAL):
            a = ob.align(mode, p, t, *s)
            assert ob.score_only(mode, p, t, *s) == (a.score, a.end_i, a.end_j)


def test_checkpointed_traceback_equals_full_matrix_oracle():
    """orc_align_ckpt (the config-4-sized oracle) == orc_align: every field and the op list, with block heights down to 1 row."""
    import random
    rng = random.Random(4)
    cases = []
    for _ in range(150):
        alpha = rng.choice([b"ACGT", b"AC", b"A", b"ACGTN"])
        m, n = rng.randint(1, 70), rng.randint(1, 90)
        p = bytes(rng.choice(alpha) for _ in range(m)); t = bytes(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.4:                                    # tandem repeats: ties everywhere
            u = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 5)))
            p = (u * 40)[:m]; t = (u * 40)[:n]
        cases.append((p, t))
    for s in ((1, -1, -1), (2, -3, -4), (5, -4, -16), (0, 0, 0), (3, 1, -2)):
        for p, t in cases[::3] if s != (1, -1, -1) else cases:
            for mode in (0, 1):
                a = ob.align(mode, p, t, *s)
                for ck in (1, 3, 16, 1024):
                    b = ob.align_ckpt(mode, p, t, *s, ck=ck)
                    assert (a.score, a.end_i, a.end_j, a.start_i, a.start_j, a.overlap, a.ops) == \
                           (b.score, b.end_i, b.end_j, b.start_i, b.start_j, b.overlap, b.ops), (mode, s, ck, p, t)


@pytest.mark.parametrize("mode,name,header", [(1, "c4_local.txt.gz", "Highest local alignment score:"), (0, "c4_global.txt.gz", "Longest overlap:")])
def test_checkpointed_oracle_reproduces_reference_file_at_config4_size(mode, name, header):
    """Pins the config-4-sized oracle: its score, CIGAR and MD:Z for the seed-482 100 kb x 100 kb pair equal, byte for byte, the file
    the UNMODIFIED reference binary wrote for that pair (tests/golden/make_golden_c4.py, ~50 GB of RAM and 3 min per mode)."""
    import gzip
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    from __graft_entry__ import load_package
    load_package()
    from bioinformatics_algorithms_b200 import workload
    p, t = workload.config4(100_000, seed=482)
    p, t = p.tobytes(), t.tobytes()
    a = ob.align_ckpt(mode, p, t, 1, -1, -1)
    mine = "%s\npattern=%s\nreference=%s\nScore =%d\nCIGAR =%s\nMD:Z=%s\n" % (header, p.decode(), t.decode(), a.score, a.cigar, a.mdz)
    assert mine.encode() == gzip.open(path, "rb").read()


def test_empty_local_alignment():
    a = ob.align(ob.LOCAL, b"acgt", b"ACGT", 1, -1, -1)
    assert (a.score, a.cigar, a.mdz, a.n_ops if hasattr(a, "n_ops") else len(a.ops)) == (0, "", "0", 0)


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/hw2 not built")
def test_live_differential_vs_reference_binary(tmp_path):
    """Thousands of cells of fresh random input through the real hw2 and the restatement."""
    rng = random.Random(20261018)
    for it in range(150):
        k = rng.randint(1, 5)
        alpha = rng.choice([b"ACGT", b"AC", b"A", b"ACGTN", b"ACDEFGHIKLMNPQRSTVWY"])
        ps = [bytes(rng.choice(alpha) for _ in range(rng.randint(1, 120))) for _ in range(k)]
        ts = [bytes(rng.choice(alpha) for _ in range(rng.randint(1, 200))) for _ in range(k)]
        s = rng.choice([(1, -1, -1), (2, -3, -4), (5, -4, -16), (1, -2, 0), (0, 0, 0), (3, 1, -2), (-1, -1, -1)])
        for flag, mode in (("-g", ob.GLOBAL), ("-l", ob.LOCAL)):
            want = ob.run_hw2_binary(ob.REF_HW2, flag, ps, ts, *s, tmp_path)
            assert ob.render_file(mode, ps, ts, *s) == want, (flag, ps, ts, s)


@pytest.mark.skipif(not os.path.exists(ob.REF_HW3), reason="oracle/_ref/hw3 not built")
def test_affine_score_restatement_matches_hw3_centre_choice(tmp_path):
    """hw3 prints no scores; its centre choice (first line of the PHYLIP body) is the arg-max of the
    distance-stage sums (hw3.cpp:231-251), so it pins orc_affine_score on whole batches."""
    import subprocess
    rng = random.Random(3)
    for it in range(25):
        k = rng.randint(3, 6)
        base = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(20, 60)))
        seqs = []
        for _ in range(k):
            s = bytearray(base)
            for _ in range(rng.randint(0, 12)):
                pos = rng.randrange(len(s))
                r = rng.random()
                if r < 0.5:
                    s[pos] = rng.choice(b"ACGT")
                elif r < 0.75 and len(s) > 5:
                    del s[pos]
                else:
                    s.insert(pos, rng.choice(b"ACGT"))
            seqs.append(bytes(s))
        sc = rng.choice([(5, -4, -16, -4), (1, -1, -2, -1), (2, -3, -5, -2)])
        fin, fout = tmp_path / "in.fa", tmp_path / "out.phy"
        with open(fin, "wb") as f:
            for i, s in enumerate(seqs):
                f.write(b">s%d\n" % i + s + b"\n")
        subprocess.check_call([ob.REF_HW3, "-i", str(fin), "-o", str(fout), "-s", "%d:%d:%d:%d" % sc],
                              stdout=subprocess.DEVNULL)
        lines = open(fout).read().split("\n")
        centre_name = lines[1].split()[0]
        sums = [0] * k
        for i in range(k):
            for j in range(i + 1, k):
                v = ob.affine_score(seqs[i], seqs[j], *sc)
                sums[i] += v
                sums[j] += v
        best, idx = None, -1
        for i, v in enumerate(sums):      # first strict max, hw3.cpp:244-251
            if best is None or v > best:
                best, idx = v, i
        assert centre_name == "s%d" % idx, (seqs, sc, sums, lines[:3])


# ---------------- hw3 affine score: pinned by alignment-derived scores and centre choices ----------------
HW3_KAT = json.load(open(os.path.join(HERE, "golden", "hw3_kat.json")))


def parse_phy(text):
    """PHYLIP as hw3 writes it (hw3.cpp:339-357): 10-char name, then the row in space-separated blocks of 10."""
    lines = text.split("\n")
    n = int(lines[0].split()[0])
    return [(ln[:10].strip(), ln[10:].replace(" ", "")) for ln in lines[1:1 + n]]


def alignment_score(a1, a2, match, mismatch, gopen, gext):
    """Score of one alignment under hw3's model (hw3.cpp:40-84): a (mis)match column scores s; a gap run of
    length L costs Go + Ge*L (opened from V, hw3.cpp:70,77), except a run that starts the alignment, which
    comes from the borders F[i][0] / E[0][j] = Go + Ge*(L-1) (hw3.cpp:44,50)."""
    total, k, n = 0, 0, len(a1)
    while k < n:
        if a1[k] != "-" and a2[k] != "-":
            total += match if a1[k] == a2[k] else mismatch
            k += 1
            continue
        row = 0 if a1[k] == "-" else 1
        e = k
        while e < n and (a1[e] == "-" if row == 0 else a2[e] == "-") and not (a1[e] == "-" and a2[e] == "-"):
            e += 1
        L = e - k
        total += gopen + gext * (L - 1 if k == 0 else L)
        k = e
    return total


def test_affine_restatement_equals_score_of_hw3_alignments_golden():
    for c in HW3_KAT["pairs"]:
        rows = parse_phy(c["phy"])
        s1, s2 = c["seqs"]
        (_, a1), (_, a2) = rows
        assert a1.replace("-", "") == s1 and a2.replace("-", "") == s2
        want = alignment_score(a1, a2, *c["s"])
        assert ob.affine_score(s1.encode(), s2.encode(), *c["s"]) == want, c


def test_affine_restatement_reproduces_hw3_centre_golden():
    for c in HW3_KAT["stars"]:
        seqs = [s.encode() for s in c["seqs"]]
        k = len(seqs)
        sums = [0] * k
        for i in range(k):
            for j in range(i + 1, k):
                v = ob.affine_score(seqs[i], seqs[j], *c["s"])
                sums[i] += v
                sums[j] += v
        idx = max(range(k), key=lambda i: (sums[i], -i))          # first strict max, hw3.cpp:244-251
        assert parse_phy(c["phy"])[0][0] == "s%d" % idx, c


# ---------------- hw4: own NW (tie order d > u > l) + distance, pinned by trees the real hw4 wrote ----------------
HW4_KAT = json.load(open(os.path.join(HERE, "golden", "hw4_kat.json")))


def test_hw4_restatement_matches_two_sequence_trees_golden():
    """For two sequences hw4 writes "(a:h,b:h):0.0;" with h = distance / 2 (hw4.cpp:179-190)."""
    for c in HW4_KAT["pairs"]:
        a, b = c["seqs"]
        score, dist, ops = ob.hw4_nw(a.encode(), b.encode(), *c["s"])
        assert c["tree"] == "(s0:%f,s1:%f):0.0;\n" % (dist / 2.0, dist / 2.0), c
        assert len(ops) - (len(a) + len(b) - len(ops)) <= dist


@pytest.mark.skipif(not os.path.exists(ob.REF_HW4), reason="oracle/_ref/hw4 not built")
def test_hw4_restatement_live_against_reference_binary(tmp_path):
    import subprocess
    rng = random.Random(44)
    for it in range(120):
        alpha = rng.choice([b"ACGT", b"AC", b"A", b"ACGTN"])
        a = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 90)))
        b = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 90)))
        s = rng.choice([(1, -1, -1), (2, -3, -4), (1, -1, 0), (3, 1, -2), (0, 0, 0)])
        (tmp_path / "in.fa").write_bytes(b">x\n" + a + b"\n>y\n" + b + b"\n")
        subprocess.check_call([ob.REF_HW4, "-i", str(tmp_path / "in.fa"), "-t", str(tmp_path / "t.txt"), "-s", *map(str, s)])
        _, dist, _ = ob.hw4_nw(a, b, *s)
        assert (tmp_path / "t.txt").read_text() == "(x:%f,y:%f):0.0;\n" % (dist / 2.0, dist / 2.0), (a, b, s)


def test_affine_traceback_restatement_equals_hw3_alignments_golden():
    """Two-sequence outputs of the unmodified hw3: the PHYLIP rows ARE alignmentString1/2 of affine_alignment(centre, other)."""
    for c in HW3_KAT["pairs"]:
        s1, s2 = (x.encode() for x in c["seqs"])
        score, ops = ob.affine_align(s1, s2, *c["s"])
        (_, a1), (_, a2) = parse_phy(c["phy"])
        assert ob.aligned_rows(ops, s1, s2) == (a1.encode(), a2.encode()), c
        assert score == ob.affine_score(s1, s2, *c["s"]) == alignment_score(a1, a2, *c["s"])
