"""Test helper (lives under tests/ because it checks against the oracle; not collected by pytest): every kernel family once on tiny inputs (for compute-sanitizer --tool memcheck): short16 4-symbol and 8-symbol fills, both tracebacks, wide32
with stored record and with checkpoint tiles, score-only tile kernel, affine score + traceback, hw4 tie order, multi-run batches."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle_binding as ob
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
rng = random.Random(1)
rnd = lambda n, a=b"ACGT": bytes(rng.choice(a) for _ in range(n))
e = pkg.Engine(0)
pat, po, txt, to = workload.config2(96, seed=3, n_rate=0.01)
ps, ts = workload.split(pat, po), workload.split(txt, to)
ps += [rnd(rng.randint(1, 300)) for _ in range(40)] + [rnd(120, b"ACGTNRYKM"), rnd(600), rnd(40, b"AC")]
ts += [rnd(rng.randint(1, 500)) for _ in range(40)] + [rnd(300, b"ACGTN"), rnd(700), rnd(3000, b"AC")]
bad = 0
def check(mode, res, ops, s):
    global bad
    for k, (p, t) in enumerate(zip(ps, ts)):
        a = ob.align(mode, p, t, *s)
        if (int(res["score"][k]), int(res["overlap"][k]), ops[k]) != (a.score, a.overlap, a.ops):
            bad += 1; print("MISMATCH", mode, k, s)
for s in ((1, -1, -1), (2, -3, -4), (5, -4, -16)):
    for mode in (pkg.GLOBAL, pkg.LOCAL):
        res, ops = e.align_batch(mode, ps, ts, *s, want_ops=True)
        check(mode, res, ops, s)
p_, po_ = pkg.pack(ps); t_, to_ = pkg.pack(ts)
both = e.align_packed_multi([pkg.GLOBAL, pkg.LOCAL], p_, po_, t_, to_, 1, -1, -1, want_ops=True)
e.align_packed(pkg.GLOBAL, p_, po_, t_, to_, 1, -1, -1, score_only=True)
e.set_option(pkg.OPT_CKPT_BYTES, 0); e.set_option(pkg.OPT_CKPT_GROUP, 1); e.set_option(pkg.OPT_CKPT_COLS, 5)
for mode in (pkg.GLOBAL, pkg.LOCAL):
    res, ops = e.align_batch(mode, ps, ts, 1, -1, -1, want_ops=True)
    check(mode, res, ops, (1, -1, -1))
res, ops = e.align_batch(pkg.GLOBAL, ps[-5:], ts[-5:], 1, -1, -1, want_ops=True, tie_hw4=True)
sc, aops = e.affine_align(ps[:6] + ps[-2:], ts[:6] + ts[-2:], 5, -4, -16, -4)
for k, (p, t) in enumerate(zip(ps[:6] + ps[-2:], ts[:6] + ts[-2:])):
    if (int(sc[k]), aops[k]) != ob.affine_align(p, t, 5, -4, -16, -4):
        bad += 1; print("MISMATCH affine", k)
e.affine_star_scores(ps[-8:], 5, -4, -16, -4)
e.close()
print("sanitize_small: mismatches", bad)
sys.exit(1 if bad else 0)
