"""Timeline of the pipelined b2a_align_batch (B2A_TRACE=1). usage: python scripts/e2e_trace.py [pairs] [seg_pairs]"""
import os, sys, time
os.environ["B2A_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
pat_np, po_np, txt_np, to_np = workload.config2(n, seed=481)
pat = pkg.pinned_empty(len(pat_np), np.uint8); pat[:] = pat_np
txt = pkg.pinned_empty(len(txt_np), np.uint8); txt[:] = txt_np
po = pkg.pinned_empty(len(po_np), np.uint64); po[:] = po_np
to = pkg.pinned_empty(len(to_np), np.uint64); to[:] = to_np
res = pkg.pinned_empty(n, pkg.RESULT_DTYPE)
e = pkg.Engine(0)
e.set_option(pkg.OPT_SEG_PAIRS, seg)
for mode in (0, 1):
    for it in range(3):
        t0 = time.perf_counter()
        e.align_packed(mode, pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
        print(f"mode {mode} iter {it}: wall {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)
e.close()
