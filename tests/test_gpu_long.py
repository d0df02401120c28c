"""Long-pair configs of BASELINE.json at FULL size on the GPU, checked against the linear-memory oracle
(oracle/hw2_oracle.c: orc_score_only / orc_affine_score) and through size-independent properties:
  config 4: one 100 kb x 100 kb pair, local and global, score + traceback (the reference needs 49 GB for it);
  config 5: 16 x 100 kb all-vs-all, score only, linear gap (hw2 scoring) and hw3's affine scoring -- sampled pairs."""
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
    p, t = workload.config4(100_000, seed=482)
    s = (1, -1, -1)
    for mode in (pkg.LOCAL, pkg.GLOBAL):
        res, ops = eng.align_batch(mode, [p.tobytes()], [t.tobytes()], *s, want_ops=True)
        assert int(res["path"][0]) == 2                                            # the int32 banded wavefront (score > int16)
        want = ob.score_only(mode, p.tobytes(), t.tobytes(), *s)                   # linear memory, ~10 s
        assert (int(res["score"][0]), int(res["end_i"][0]), int(res["end_j"][0])) == want
        sc, si, sj = rescore(ops[0], p, t, int(res["end_i"][0]), int(res["end_j"][0]), s)
        assert (sc, si, sj) == (int(res["score"][0]), int(res["start_i"][0]), int(res["start_j"][0]))
        if mode == pkg.GLOBAL:
            assert (si, sj) == (0, 0)


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
