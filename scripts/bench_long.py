"""Long-pair configs of BASELINE.json (parity-test cases with timings, not the bench.py line):
  config 4: ONE 100 kb x 100 kb pair, local (and global), score + traceback      -> wide32 fill + traceback
  config 5: 16 sequences x 100 kb all-vs-all (120 pairs), global, score only      -> wide32 (linear gap, hw2 scoring)
                                                                                   and affine32 (hw3 scoring 5:-4:-16:-4)
usage: python scripts/bench_long.py [--len 100000] [--skip4] [--skip5]
Timing only; the traceback op list is re-scored on the host as a sanity check.  Parity against the linear-memory oracle at
these sizes is tests/test_gpu_long.py."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload

ap = argparse.ArgumentParser()
ap.add_argument("--len", type=int, default=100_000)
ap.add_argument("--check", type=int, default=0, help="(ignored; parity lives in tests/test_gpu_long.py)")
ap.add_argument("--skip4", action="store_true")
ap.add_argument("--skip5", action="store_true")
ap.add_argument("--seqs", type=int, default=16)
args = ap.parse_args()
out = {}
eng = pkg.Engine(0)


def rescore(ops, p, t, ei, ej, s):
    codes = np.frombuffer(ops, dtype=np.uint8)
    di = (codes != 0x49).astype(np.int64)      # M and D consume a pattern base
    dj = (codes != 0x44).astype(np.int64)      # M and I consume a text base
    i = ei - np.cumsum(di); j = ej - np.cumsum(dj)
    isM = codes == 0x4D
    eq = p[i[isM]] == t[j[isM]]
    score = int(eq.sum()) * s[0] + int((~eq).sum()) * s[1] + int((~isM).sum()) * s[2]
    return score, int(i[-1]) if len(i) else ei, int(j[-1]) if len(j) else ej


if not args.skip4:
    p, t = workload.config4(args.len, seed=482)
    s = (1, -1, -1)
    for mode, name in ((pkg.LOCAL, "local"), (pkg.GLOBAL, "global")):
        pat, po = pkg.pack([p.tobytes()]); txt, to = pkg.pack([t.tobytes()])
        eng.upload(mode, pat, po, txt, to, *s, want_ops=True)
        eng.run()
        best = min((eng.run(), eng.times())[1] for _ in range(3))
        res = eng.download(1)
        cells = len(p) * len(t)
        r = {"m": len(p), "n": len(t), "fill_ms": best[0], "traceback_ms": best[1], "total_ms": best[2],
             "gcups_fill": cells / best[0] / 1e6, "gcups_total": cells / best[2] / 1e6,
             "score": int(res["score"][0]), "n_ops": int(res["n_ops"][0]), "path": int(res["path"][0])}
        words, off = eng.copy_ops(1)
        ops = pkg.unpack_ops(words, off, 0, res["n_ops"][0])
        sc, si, sj = rescore(ops, p, t, int(res["end_i"][0]), int(res["end_j"][0]), s)
        r["ops_rescore_ok"] = bool(sc == r["score"] and (si, sj) == (int(res["start_i"][0]), int(res["start_j"][0])))
        out["config4_" + name] = r
        print("config4", name, json.dumps(r), flush=True)

if not args.skip5:
    rng = np.random.default_rng(4830)
    sets = {"iid": workload.config5(args.seqs, args.len, seed=483)}
    # shape of Multiple_Sequence_Alignment/input16100000.fasta: tandem repeats of period 5..1001, lengths L, L+10, L+20, L+100
    tand = []
    for k in range(args.seqs):
        per = [5, 7, 20, 1000, 1001][k % 5]
        unit = workload.ACGT[rng.integers(0, 4, size=per, dtype=np.uint8)]
        L = args.len + [0, 0, 0, 0, 0, 0, 0, 10, 20, 100, 100, 100, 100, 100, 100, 100][k % 16]
        tand.append(np.tile(unit, L // per + 1)[:L].copy())
    sets["tandem"] = tand
    for sname, seqs in sets.items():
        sb = [x.tobytes() for x in seqs]
        ij = [(i, j) for i in range(len(sb)) for j in range(i + 1, len(sb))]
        cells = sum(len(sb[i]) * len(sb[j]) for i, j in ij)
        # linear gap (hw2 scoring), score only
        pat, po = pkg.pack([sb[i] for i, _ in ij]); txt, to = pkg.pack([sb[j] for _, j in ij])
        eng.upload(pkg.GLOBAL, pat, po, txt, to, 1, -1, -1, score_only=True)
        eng.run()
        best = min((eng.run(), eng.times())[1] for _ in range(2))
        res = eng.download(len(ij))
        r = {"pairs": len(ij), "cells": cells, "total_ms": best[2], "gcups": cells / best[2] / 1e6}
        out[f"config5_{sname}_linear"] = r
        print("config5", sname, "linear", json.dumps(r), flush=True)
        # hw3 affine scoring
        t0 = time.perf_counter()
        ps, sums, centre = eng.affine_star_scores(sb, 5, -4, -16, -4)
        wall = (time.perf_counter() - t0) * 1e3
        ms = eng.times()[2]
        r = {"pairs": len(ij), "cells": cells, "kernel_ms": ms, "wall_ms": wall, "gcups": cells / ms / 1e6, "centre": int(centre)}
        out[f"config5_{sname}_affine"] = r
        print("config5", sname, "affine", json.dumps(r), flush=True)
eng.close()
print(json.dumps(out))
