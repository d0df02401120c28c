"""End-to-end time of ONE rank's share of a strong-scaled batch (125 k pairs, both modes, one upload) under different segment schedules,
with the engine's per-segment timeline (B2A_TRACE=1).  usage: python scripts/small_batch_exp.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
pat_np, po_np, txt_np, to_np = workload.config2(n, seed=481)
def pin(a):
    o = pkg.pinned_empty(len(a), a.dtype); o[:] = a; return o
pat, po, txt, to = pin(pat_np), pin(po_np), pin(txt_np), pin(to_np)
res = [pkg.pinned_empty(n, pkg.RESULT_DTYPE) for _ in range(2)]
W = 7104
for name, first, mx, lanes in (("library default", 0, 0, 0), ("16k doubling to 128k, 8 lanes", 16384, 131072, 8), ("16k doubling to 32k, 8 lanes", 16384, 32768, 8),
                               ("8k doubling to 32k, 8 lanes", 8192, 32768, 8), ("16k flat, 8 lanes", 16384, 16384, 8), ("16k doubling to 64k, 8 lanes", 16384, 65536, 8),
                               ("24k flat, 8 lanes", 24576, 24576, 8), ("16k doubling to 128k, 2 lanes", 16384, 131072, 2)):
    os.environ["B2A_TRACE"] = "0"
    e = pkg.Engine(0)
    if first:
        e.set_option(pkg.OPT_SEG_FIRST, first); e.set_option(pkg.OPT_SEG_PAIRS, mx)
    if lanes:
        e.set_option(pkg.OPT_LANES, lanes)
    for _ in range(2):
        e.align_packed_multi([0, 1], pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        e.align_packed_multi([0, 1], pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:30s}: e2e ms min {min(ts):.2f} median {np.median(ts):.2f}; launches {e.stats()['launches']}", flush=True)
    e.close()
for name, first, mx in (("library default", 0, 0),):
    os.environ["B2A_TRACE"] = "1"
    e = pkg.Engine(0)
    if first:
        e.set_option(pkg.OPT_SEG_FIRST, first); e.set_option(pkg.OPT_SEG_PAIRS, mx)
    print("---- trace:", name, flush=True)
    for _ in range(3):
        e.align_packed_multi([0, 1], pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
    e.close()
