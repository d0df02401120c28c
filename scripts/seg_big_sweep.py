"""End-to-end time of b2a_align_batch_multi (config 2, both modes) with segments LARGER than the default 128 k pairs / 8 GB of record:
a traceback kernel needs ~1.3 ms however few pairs it walks (one resident wave = 7104 warps = 227 k pairs), so larger segments amortise it.
usage: python scripts/seg_big_sweep.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pat_np, po_np, txt_np, to_np = workload.config2(n, seed=481)
def pin(a):
    o = pkg.pinned_empty(len(a), a.dtype); o[:] = a; return o
pat, po, txt, to = pin(pat_np), pin(po_np), pin(txt_np), pin(to_np)
res = [pkg.pinned_empty(n, pkg.RESULT_DTYPE) for _ in range(2)]
W = 7104 * 32
for first, mx, gb, lanes in ((16384, 131072, 8, 4), (16384, W, 16, 4), (16384, W, 16, 3), (32768, W, 16, 4), (16384, 2 * W, 32, 2), (16384, 2 * W, 32, 3),
                             (16384, 3 * W // 2, 24, 3), (16384, 262144, 16, 4)):
    e = pkg.Engine(0)
    e.set_option(pkg.OPT_SEG_FIRST, first); e.set_option(pkg.OPT_SEG_PAIRS, mx); e.set_option(pkg.OPT_SEG_BYTES, gb << 30); e.set_option(pkg.OPT_LANES, lanes)
    e.align_packed_multi([0, 1], pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        e.align_packed_multi([0, 1], pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"first {first:6d} max {mx:6d} record {gb:2d} GB lanes {lanes}: e2e ms {min(ts):.2f} (min) {np.median(ts):.2f} (median); launches {e.stats()['launches']}", flush=True)
    e.close()
