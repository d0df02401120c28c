"""CPU tests of the short16 HBM record + the traceback walkers the CUDA kernel runs.

tests/hostmodel.cpp encodes the record from a plain DP matrix (checking the delta-range lemma and
the ring-arithmetic word encoding on the way) and runs the shared walkers of b2a_format.h; results
must equal the oracle's (score, coordinates, overlap, op list) exactly.
"""
import ctypes as C
import os
import random
import subprocess

import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")


class PairResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("end_i", C.c_uint32), ("end_j", C.c_uint32), ("start_i", C.c_uint32),
                ("start_j", C.c_uint32), ("overlap", C.c_int32), ("n_ops", C.c_uint32), ("path", C.c_uint32)]


def hostmodel():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libhostmodel.so")
    srcs = [os.path.join(HERE, "hostmodel.cpp"), os.path.join(ROOT, "bioinformatics-algorithms_b200", "csrc", "b2a_format.h"),
            os.path.join(ROOT, "bioinformatics-algorithms_b200", "csrc", "fasta_hw2.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-o", so, srcs[0], "-lpthread"])
    lib = C.CDLL(so)
    lib.hm_record_chunks.restype = C.c_uint64
    return lib


def unpack_ops(words, n_ops):
    return bytes(b"MDI"[(words[t // 16] >> (2 * (t % 16))) & 3] for t in range(n_ops))


def model_pairpair(lib, mode, pa, ta, pb, tb, s):
    m, n = len(pa), len(ta)
    assert (len(pb), len(tb)) == (m, n)
    plan = (C.c_int * 3)()
    if not lib.hm_plan(mode, m, n, s[0], s[1], s[2], plan):
        return None
    K, R, bias = plan[0], plan[1], plan[2]
    res = (PairResult * 2)()
    nw = (m + n + 15) // 16 + 1
    oa, obuf = (C.c_uint32 * nw)(), (C.c_uint32 * nw)()
    rc = lib.hm_run_pairpair(mode, K, R, pa, pb, m, ta, tb, n, s[0], s[1], s[2], bias, pa[0], res, oa, obuf, None, None)
    assert rc == 0, f"hostmodel rc={rc} (K={K} R={R} bias={bias})"
    return [(res[0], unpack_ops(oa, res[0].n_ops)), (res[1], unpack_ops(obuf, res[1].n_ops))], (K, R, bias)


def check(lib, mode, pa, ta, pb, tb, s):
    got = model_pairpair(lib, mode, pa, ta, pb, tb, s)
    if got is None:
        return False
    for (r, ops), (p, t) in zip(got[0], ((pa, ta), (pb, tb))):
        a = ob.align(mode, p, t, *s)
        assert (r.score, r.end_i, r.end_j, r.start_i, r.start_j, r.overlap, ops) == \
               (a.score, a.end_i, a.end_j, a.start_i, a.start_j, a.overlap, a.ops), (mode, p, t, s, got[1], a)
    return True


def rnd(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def mutate(rng, s, psub=0.08, pindel=0.02):
    out = bytearray()
    for ch in s:
        r = rng.random()
        if r < pindel:
            continue
        if r < 2 * pindel:
            out.append(rng.choice(b"ACGT"))
        out.append(rng.choice(b"ACGT") if rng.random() < psub else ch)
    return bytes(out)


SCORINGS = [(1, -1, -1), (2, -3, -4), (5, -4, -16), (3, -2, -1), (1, -1, -2), (2, -1, -1), (1, 0, 0), (4, -6, -3), (0, 0, 0)]


def test_plan_picks_expected_widths():
    lib = hostmodel()
    assert lib.hm_check_rmagic() == 1
    assert lib.hm_delta_bits(1, -1, -1) == 2
    assert lib.hm_delta_bits(2, -3, -4) == 4
    assert lib.hm_delta_bits(5, -4, -16) == 8
    assert lib.hm_delta_bits(1, -1, 1) == 0      # positive "gap" reward: not representable, wide path
    plan = (C.c_int * 3)()
    assert lib.hm_plan(0, 150, 1000, 1, -1, -1, plan) and (plan[0], plan[1]) == (2, 5)
    assert lib.hm_plan(0, 150, 1000, 5, -4, -16, plan) and plan[0] == 8
    assert lib.hm_plan(0, 300, 1000, 1, -1, -1, plan) and plan[1] == 10      # 257..320 rows: 10 rows per lane
    assert lib.hm_plan(0, 400, 1000, 1, -1, -1, plan) and plan[1] == 16
    assert not lib.hm_plan(0, 513, 1000, 1, -1, -1, plan)    # > 512 rows: int32 family
    assert not lib.hm_plan(1, 150, 100000, 1, -1, -1, plan)  # too many columns


def test_walkers_match_oracle_random_shapes():
    lib = hostmodel()
    rng = random.Random(4811)
    done = 0
    for it in range(260):
        m, n = rng.randint(1, 70), rng.randint(1, 120)
        kind = rng.random()
        pairs = []
        for _ in range(2):
            if kind < 0.3:   # tie stress
                u1, u2 = rnd(rng, rng.randint(1, 3), b"AC"), rnd(rng, rng.randint(1, 3), b"AC")
                p, t = (u1 * 200)[:m], (u2 * 200)[:n]
            elif kind < 0.7:
                t = rnd(rng, n)
                src = mutate(rng, t[rng.randint(0, max(0, n - m)):])
                p = (src + rnd(rng, m))[:m]
            else:
                p, t = rnd(rng, m), rnd(rng, n)
            pairs.append((p, t))
        s = rng.choice(SCORINGS)
        for mode in (ob.GLOBAL, ob.LOCAL):
            done += check(lib, mode, pairs[0][0], pairs[0][1], pairs[1][0], pairs[1][1], s)
    assert done > 400


def test_walkers_match_oracle_config2_shape():
    lib = hostmodel()
    rng = random.Random(482)
    for it in range(6):
        pairs = []
        for _ in range(2):
            t = rnd(rng, 1000)
            off = rng.randint(0, 850)
            p = (mutate(rng, t[off:off + 150], 0.05, 0.005) + rnd(rng, 150))[:150]
            pairs.append((p, t))
        for s in ((1, -1, -1), (2, -3, -4)):
            for mode in (ob.GLOBAL, ob.LOCAL):
                assert check(lib, mode, pairs[0][0], pairs[0][1], pairs[1][0], pairs[1][1], s)


def test_walkers_rows_up_to_256_and_homopolymers():
    lib = hostmodel()
    for m, n in ((256, 300), (255, 40), (33, 33), (32, 1), (1, 32), (1, 1), (97, 257)):
        for p, t in ((b"A" * m, b"A" * n), (b"AC" * m, b"CA" * n), (b"A" * m, b"C" * n)):
            for s in ((1, -1, -1), (2, -3, -4), (1, 0, 0)):
                for mode in (ob.GLOBAL, ob.LOCAL):
                    assert check(lib, mode, p[:m], t[:n], p[:m][::-1], t[:n][::-1], s)


# ---------------- wide32 record (int32, bands of 128 rows, K in {2,4,8,16,32}) ----------------
def check_wide(lib, mode, p, t, s, K=0):
    res = PairResult()
    nw = (len(p) + len(t) + 15) // 16 + 1
    ops = (C.c_uint32 * nw)()
    rc = lib.hm_run_wide(mode, K, p, len(p), t, len(t), s[0], s[1], s[2], C.byref(res), ops)
    assert rc == 0, f"hostmodel wide rc={rc}"
    a = ob.align(mode, p, t, *s)
    assert (res.score, res.end_i, res.end_j, res.start_i, res.start_j, res.overlap, unpack_ops(ops, res.n_ops)) == \
           (a.score, a.end_i, a.end_j, a.start_i, a.start_j, a.overlap, a.ops), (mode, p, t, s, K, a)


WIDE_SCORINGS = SCORINGS + [(100, -100, -200), (1, -1, 1), (-1, -2, -1), (-3, -3, -3), (7, 9, -2), (300, -200, -5000)]


def test_wide_delta_bits():
    lib = hostmodel()
    assert lib.hm_delta_bits_wide(1, -1, -1) == 2
    assert lib.hm_delta_bits_wide(5, -4, -16) == 8
    assert lib.hm_delta_bits_wide(100, -100, -200) == 16
    assert lib.hm_delta_bits_wide(300, -200, -50000) == 32
    assert lib.hm_delta_bits_wide(1, -1, 1) == 32       # lemma does not apply: raw deltas
    assert lib.hm_delta_bits_wide(-1, -2, -1) == 32


def test_wide_walkers_match_oracle():
    lib = hostmodel()
    rng = random.Random(777)
    for it in range(160):
        m, n = rng.randint(1, 300), rng.randint(1, 200)
        alpha = rng.choice([b"ACGT", b"AC", b"ACDEFGHIKLMNPQRSTVWYacgt"])
        if rng.random() < 0.5:
            t = rnd(rng, n, alpha)
            p = (mutate(rng, t[rng.randint(0, max(0, n - 1)):]) + rnd(rng, m, alpha))[:m]
        else:
            p, t = rnd(rng, m, alpha), rnd(rng, n, alpha)
        s = rng.choice(WIDE_SCORINGS)
        for mode in (ob.GLOBAL, ob.LOCAL):
            check_wide(lib, mode, p, t, s)


def test_wide_every_k_on_the_same_input():
    lib = hostmodel()
    rng = random.Random(5)
    t = rnd(rng, 333)
    p = mutate(rng, t[40:300])
    for K in (2, 4, 8, 16, 32):
        for mode in (ob.GLOBAL, ob.LOCAL):
            check_wide(lib, mode, p, t, (1, -1, -1), K)


def test_walker_with_hw4_tie_order_matches_hw4_oracle():
    """walker bit 4 = hw4's tie order d > u > l (hw4.cpp:37-46): ops must equal hw4's own needleman_wunsch, and differ from hw2's
    order on inputs where 'u' and 'l' tie (the delta record itself is tie-order agnostic)."""
    lib = hostmodel()
    rng = random.Random(17)
    differs = 0
    try:
        for it in range(120):
            alpha = rng.choice([b"ACGT", b"AC", b"A"])
            m, n = rng.randint(1, 60), rng.randint(1, 80)
            pa, pb = rnd(rng, m, alpha), rnd(rng, m, alpha)
            ta, tb = rnd(rng, n, alpha), rnd(rng, n, alpha)
            s = rng.choice([(1, -1, -1), (2, -3, -4), (1, 0, 0)])
            for opt in (4, 4 | 2):
                lib.hm_set_opt(opt)
                got = model_pairpair(lib, ob.GLOBAL, pa, ta, pb, tb, s)
                if got is None:
                    continue
                for (r, ops), (p, t) in zip(got[0], ((pa, ta), (pb, tb))):
                    score, dist, want = ob.hw4_nw(p, t, *s)
                    assert (r.score, ops) == (score, want), (p, t, s)
                    differs += ops != ob.align(ob.GLOBAL, p, t, *s).ops
    finally:
        lib.hm_set_opt(3)
    assert differs > 0


def test_walkers_match_oracle_long_patterns():
    """rows-per-lane 10, 12, 16 (patterns of 257..512 bases): record layout + walkers against the oracle on the CPU"""
    lib = hostmodel()
    rng = random.Random(4830)
    for m, n in ((257, 300), (300, 420), (384, 200), (400, 500), (512, 130)):
        t1, t2 = rnd(rng, n), rnd(rng, n)
        p1 = (mutate(rng, t1) + rnd(rng, m))[:m]
        p2 = (mutate(rng, t2) + rnd(rng, m))[:m]
        for mode in (ob.GLOBAL, ob.LOCAL):
            assert check(lib, mode, p1, t1, p2, t2, (1, -1, -1))


def read_fasta_hw2(raw: bytes):
    """readFasta, hw2.cpp:25-57, restated line by line: getline, trailing whitespace stripped (:38), blank lines skipped (:39), a '>' line
    flushes the current record only if it is non-empty (:41-46), everything else is appended (:48), last record flushed if non-empty (:52)."""
    seqs, cur = [], b""
    for line in raw.split(b"\n"):
        line = line.rstrip(b" \t\r\n\v\f")
        if not line:
            continue
        if line[:1] == b">":
            if cur:
                seqs.append(cur); cur = b""
        else:
            cur += line
    if cur:
        seqs.append(cur)
    return seqs


def test_parallel_fasta_loader_equals_serial_reader_semantics(tmp_path):
    """bin/hw2's multi-threaded FASTA loader (csrc/fasta_hw2.h) on messy files, with 1..8 threads forced even on tiny inputs so that chunk
    borders fall inside records, between headers, on blank lines and on CRLF lines."""
    import random
    lib = hostmodel()
    lib.hm_load_fasta.restype = C.c_int64
    rng = random.Random(11)
    files = [b"", b"\n\n", b">only header\n", b"ACGT", b"ACGT\n>h\n>h2\n\nTT\r\nGG  \n>x\n", b">a\nAC\n>b\n\n>c\nGT\n"]
    for _ in range(120):
        parts = []
        for _ in range(rng.randint(0, 40)):
            r = rng.random()
            if r < 0.30:
                parts.append(b">" + bytes(rng.choice(b"abc >") for _ in range(rng.randint(0, 6))))
            elif r < 0.40:
                parts.append(rng.choice([b"", b"  ", b"\r", b"\t"]))
            else:
                parts.append(bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(1, 30))) + rng.choice([b"", b" ", b"\r", b" \t\r"]))
        files.append(b"\n".join(parts) + rng.choice([b"", b"\n", b"\n\n"]))
    big = b"".join(b">r%d\n%s\n" % (k, bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 200)))) for k in range(20000))
    files.append(big)
    path = str(tmp_path / "x.fa").encode()
    for raw in files:
        with open(path, "wb") as f:
            f.write(raw)
        want = read_fasta_hw2(raw)
        for threads in (1, 2, 3, 5, 8):
            bytes_buf = (C.c_uint8 * (len(raw) + 1))()
            off_buf = (C.c_uint64 * (raw.count(b">") + 3))()
            n = lib.hm_load_fasta(path, threads, bytes_buf, len(raw) + 1, off_buf, len(off_buf))
            assert n == len(want), (threads, raw[:80])
            got = [bytes(bytes_buf[off_buf[k]:off_buf[k + 1]]) for k in range(n)]
            assert got == want, (threads, raw[:80])
    assert lib.hm_load_fasta(b"/nonexistent/file.fa", 2, None, 0, None, 0) == -1
