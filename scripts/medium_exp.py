"""Batches just outside the s16x2 path (pattern > 256 rows): wide32 throughput on many medium pairs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
e = pkg.Engine(0)
for m, n, pairs in ((300, 1000, 200000), (500, 2000, 50000), (256, 1000, 200000)):
    pat, po, txt, to = workload.config2(pairs, seed=5, m=m, n=n)
    for mode in (0, 1):
        e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
        e.run()
        best = min((e.run(), e.times())[1] for _ in range(2))
        res = e.download(pairs)
        print(f"{m}x{n} x {pairs} mode {mode}: path {set(map(int, res['path']))} fill {best[0]:.2f} tb {best[1]:.2f} ms -> {pairs * m * n / best[2] / 1e6:.0f} GCUPS", flush=True)
e.close()
