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
for _once in (0,):
  for name, first, mx, lanes in (("library default", 0, 0, 0), ("16k doubling to 96k, 2 lanes", 16384, 98304, 2), ("uniform 2 waves, 2 lanes", 2 * W, 2 * W, 2),
                                 ("uniform 2 waves, 4 lanes", 2 * W, 2 * W, 4), ("uniform 2 waves, 8 lanes", 2 * W, 2 * W, 8), ("uniform 1 wave, 8 lanes", W, W, 8),
                                 ("uniform 4 waves, 8 lanes", 4 * W, 4 * W, 8), ("16k doubling, 8 lanes", 16384, 98304, 8), ("2 segments, 2 lanes", 62500, 62500, 2)):
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
