"""Ragged batch (every pair its own shape): throughput with and without mixed-shape pair-pairs (B2A_NO_MIX=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
rng = np.random.default_rng(7)
n_pairs = 200000
ms = rng.integers(100, 201, size=n_pairs); ns = rng.integers(800, 1201, size=n_pairs)
po = np.zeros(n_pairs + 1, dtype=np.uint64); to = np.zeros(n_pairs + 1, dtype=np.uint64)
np.cumsum(ms, out=po[1:]); np.cumsum(ns, out=to[1:])
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
pat = acgt[rng.integers(0, 4, size=int(po[-1]), dtype=np.uint8)]; txt = acgt[rng.integers(0, 4, size=int(to[-1]), dtype=np.uint8)]
cells = float((ms * ns).sum())
e = pkg.Engine(0)
for mode in (0, 1):
    e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
    e.run()
    best = min((e.run(), e.times())[1] for _ in range(2))
    res = e.download(n_pairs)
    print(f"mode {mode}: fill {best[0]:.2f} tb {best[1]:.2f} ms -> {cells / best[2] / 1e6:.0f} GCUPS; checksum {int(res['score'].astype(np.int64).sum())} {int(res['n_ops'].astype(np.int64).sum())} {int(res['overlap'].astype(np.int64).sum())}", flush=True)
e.close()
