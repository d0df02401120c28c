"""How long does one 128-row band take per 32-column block, alone and chained? (wide32 fill, n = 100 k)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
rng = np.random.default_rng(1)
n = 100000
t = workload.ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
eng = pkg.Engine(0)
for mode in (0,):
    for m in (int(sys.argv[1]) if len(sys.argv) > 1 else 12800,):
        p = workload.mutate(rng, t, 0.08, 0.01, 0.01)[:m] if m < n else workload.mutate(rng, t, 0.08, 0.01, 0.01)
        pat, po = pkg.pack([p.tobytes()]); txt, to = pkg.pack([t.tobytes()])
        for so in (False,):
            eng.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=not so, score_only=so)
            eng.run()
            best = min((eng.run(), eng.times())[1][0] for _ in range(1))
            bands = (len(p) + 127) // 128
            blocks = (n + 63) // 32 + 3 * (bands - 1)
            print(f"mode {mode} m {len(p):6d} bands {bands:4d} score_only {so}: fill {best:8.3f} ms = {best * 1e3 / blocks:6.3f} us per block-slot ({best * 1e3 / ((n + 63) // 32):6.3f} us per own block)", flush=True)
eng.close()
