"""Long-pair configs of BASELINE.json at FULL size on the GPU:
  config 4: one 100 kb x 100 kb pair, local and global, score + traceback -- the drop-in binary's output file against the file the
            UNMODIFIED reference binary wrote for the same pair (tests/golden/c4_*.txt.gz, made by make_golden_c4.py with 50 GB of RAM),
            the op list against the oracle's row-checkpointed traceback (orc_align_ckpt) byte for byte, score + end cell against the
            linear-memory oracle;
  config 5: the reference's own 16 x 100 kb input (tests/golden/input16100000.fasta.gz): ALL 120 linear and ALL 120 affine scores
            against the linear-memory oracle (orc_score_only / orc_affine_score) on every host core, plus the synthetic iid variant sampled."""
import gzip
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_binding as ob
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
pkg = load_package()
from bioinformatics_algorithms_b200 import workload  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    e = pkg.Engine(0)
    yield e
    e.close()


def rescore(ops, p, t, ei, ej, s):
    codes = np.frombuffer(ops, dtype=np.uint8)
    di = (codes != 0x49).astype(np.int64)          # M and D consume a pattern base
    dj = (codes != 0x44).astype(np.int64)          # M and I consume a text base
    i = ei - np.cumsum(di); j = ej - np.cumsum(dj)
    isM = codes == 0x4D
    eq = p[i[isM]] == t[j[isM]]
    return int(eq.sum()) * s[0] + int((~eq).sum()) * s[1] + int((~isM).sum()) * s[2], int(i[-1]), int(j[-1])


def test_config4_single_100kb_pair(eng):
    """served by the int32 banded wavefront (scores beyond int16); the op list consumes exactly the reported cells and re-scores to the
    reported score (the comparison with the oracle, op for op, is test_config4_ops_equal_checkpointed_oracle)"""
    p, t = workload.config4(100_000, seed=482)
    s = (1, -1, -1)
    for mode in (pkg.LOCAL, pkg.GLOBAL):
        res, ops = eng.align_batch(mode, [p.tobytes()], [t.tobytes()], *s, want_ops=True)
        assert int(res["path"][0]) == 2
        sc, si, sj = rescore(ops[0], p, t, int(res["end_i"][0]), int(res["end_j"][0]), s)
        assert (sc, si, sj) == (int(res["score"][0]), int(res["start_i"][0]), int(res["start_j"][0]))
        if mode == pkg.GLOBAL:
            assert (si, sj) == (0, 0)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("flag,name", [("-l", "c4_local.txt.gz"), ("-g", "c4_global.txt.gz")])
def test_config4_cli_file_equals_reference_binary_output(eng, tmp_path, flag, name):
    """bin/hw2 on the seed-482 100 kb pair writes byte for byte the file the unmodified hw2 wrote (CIGAR and MD:Z of the reference's own
    tie-broken path, hw2.cpp:145-153 / :214-222)."""
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    p, t = workload.config4(100_000, seed=482)
    for fn, seq in (("p.fa", p), ("t.fa", t)):
        (tmp_path / fn).write_bytes(b">s\n" + seq.tobytes() + b"\n")
    subprocess.check_call([pkg.HW2_BIN, flag, "-p", "p.fa", "-t", "t.fa", "-o", "out.txt", "-s", "1", "-1", "-1"], cwd=tmp_path)
    assert (tmp_path / "out.txt").read_bytes() == gzip.open(path, "rb").read()


def test_config4_ops_equal_checkpointed_oracle(eng):
    """The whole 100 kb op list (not a re-scoring) against the oracle's row-checkpointed traceback, local and global."""
    p, t = workload.config4(100_000, seed=482)
    ob.lib()
    with ThreadPoolExecutor(max_workers=2) as ex:                               # ~1 min each on one core; ctypes releases the GIL
        want = dict(zip((pkg.LOCAL, pkg.GLOBAL), ex.map(lambda mode: ob.align_ckpt(mode, p.tobytes(), t.tobytes(), 1, -1, -1), (pkg.LOCAL, pkg.GLOBAL))))
    for mode in (pkg.LOCAL, pkg.GLOBAL):
        res, ops = eng.align_batch(mode, [p.tobytes()], [t.tobytes()], 1, -1, -1, want_ops=True)
        a = want[mode]
        got = (int(res["score"][0]), int(res["end_i"][0]), int(res["end_j"][0]), int(res["start_i"][0]), int(res["start_j"][0]), int(res["overlap"][0]))
        assert got == (a.score, a.end_i, a.end_j, a.start_i, a.start_j, a.overlap)
        assert ops[0] == a.ops


def test_config4_checkpointed_walk_equals_stored_record_walk(eng):
    """The 100 kb pair again through checkpointed recomputation (score pass that keeps every 16th band's bottom row and every 8192-th column,
    then 2048 x 8192 tiles re-filled along the path): the same record and the same 101 138 ops as the walk over the full 5 GB record."""
    p, t = workload.config4(100_000, seed=482)
    e = pkg.Engine(0)
    try:
        for mode in (pkg.LOCAL, pkg.GLOBAL):
            res, ops = eng.align_batch(mode, [p.tobytes()], [t.tobytes()], 1, -1, -1, want_ops=True)
            e.set_option(pkg.OPT_CKPT_BYTES, 1 << 30); e.set_option(pkg.OPT_CKPT_GROUP, 100 << 20)       # groups of 16 bands, tiles of 8192 columns
            res2, ops2 = e.align_batch(mode, [p.tobytes()], [t.tobytes()], 1, -1, -1, want_ops=True)
            assert res[0] == res2[0] and ops[0] == ops2[0]
    finally:
        e.close()


def test_config5_shipped_input_all_120_pairs(eng):
    """Every pair of the reference's own input16100000.fasta (tandem repeats: extreme ties), linear and affine, against the oracle."""
    seqs = [x for _, x in workload.config5_shipped()]
    assert len(seqs) == 16 and sorted(map(len, seqs)) == [100000] * 7 + [100010, 100020] + [100100] * 7
    ij = [(i, j) for i in range(16) for j in range(i + 1, 16)]
    step = int(os.environ.get("B2A_C5_STRIDE", "1"))               # builder runs may subsample; the default checks all 120
    pat, po = pkg.pack([seqs[i] for i, _ in ij]); txt, to = pkg.pack([seqs[j] for _, j in ij])
    res = eng.align_packed(pkg.GLOBAL, pat, po, txt, to, 1, -1, -1, score_only=True)
    ps, sums, centre = eng.affine_star_scores(seqs, 5, -4, -16, -4)
    ob.lib()
    ks = list(range(0, len(ij), step))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:             # ctypes releases the GIL: one oracle call per core
        lin = list(ex.map(lambda k: ob.score_only(pkg.GLOBAL, seqs[ij[k][0]], seqs[ij[k][1]], 1, -1, -1)[0], ks))
        aff = list(ex.map(lambda k: ob.affine_score(seqs[ij[k][0]], seqs[ij[k][1]], 5, -4, -16, -4), ks))
    assert [int(res["score"][k]) for k in ks] == lin
    assert [int(ps[k]) for k in ks] == aff
    want = np.zeros(16, dtype=np.int64)
    for k, (i, j) in enumerate(ij):
        want[i] += int(ps[k]); want[j] += int(ps[k])
    assert list(map(int, sums)) == list(map(int, want)) and centre == int(np.argmax(want))


def test_config5_all_vs_all_100kb_sampled(eng):
    seqs = [x.tobytes() for x in workload.config5(16, 100_000, seed=483)]
    ij = [(i, j) for i in range(16) for j in range(i + 1, 16)]
    pat, po = pkg.pack([seqs[i] for i, _ in ij]); txt, to = pkg.pack([seqs[j] for _, j in ij])
    res = eng.align_packed(pkg.GLOBAL, pat, po, txt, to, 1, -1, -1, score_only=True)
    ps, sums, centre = eng.affine_star_scores(seqs, 5, -4, -16, -4)
    for k in (0, 77):
        i, j = ij[k]
        assert int(res["score"][k]) == ob.score_only(pkg.GLOBAL, seqs[i], seqs[j], 1, -1, -1)[0]
    i, j = ij[53]
    assert int(ps[53]) == ob.affine_score(seqs[i], seqs[j], 5, -4, -16, -4)
    want = np.zeros(16, dtype=np.int64)
    for k, (i, j) in enumerate(ij):
        want[i] += int(ps[k]); want[j] += int(ps[k])
    assert list(map(int, sums)) == list(map(int, want)) and centre == int(np.argmax(want))
