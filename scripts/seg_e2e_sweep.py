"""End-to-end time of b2a_align_batch (pinned host buffers, 1 M pairs) for a few segment schedules."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = 1000000
pat_np, po_np, txt_np, to_np = workload.config2(n, seed=481)
pat = pkg.pinned_empty(len(pat_np), np.uint8); pat[:] = pat_np
txt = pkg.pinned_empty(len(txt_np), np.uint8); txt[:] = txt_np
po = pkg.pinned_empty(len(po_np), np.uint64); po[:] = po_np
to = pkg.pinned_empty(len(to_np), np.uint64); to[:] = to_np
res = pkg.pinned_empty(n, pkg.RESULT_DTYPE)
e = pkg.Engine(0)
for lanes, first, mx in ((1, 16384, 98304), (2, 16384, 98304), (2, 16384, 65536), (2, 16384, 131072), (2, 32768, 131072), (2, 8192, 65536), (2, 16384, 49152)):
    e.set_option(pkg.OPT_LANES, lanes); e.set_option(pkg.OPT_SEG_PAIRS, mx); e.set_option(pkg.OPT_SEG_FIRST, first); e.set_option(pkg.OPT_SEG_BYTES, 1 << 40)
    tot = 0.0
    for mode in (0, 1):
        e.align_packed(mode, pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
        w = []
        for _ in range(4):
            t0 = time.perf_counter()
            e.align_packed(mode, pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
            w.append((time.perf_counter() - t0) * 1e3)
        tot += min(w)
    print(f"lanes {lanes} first {first:7d} max {mx:7d}: NW+SW e2e {tot:6.2f} ms", flush=True)
e.close()
