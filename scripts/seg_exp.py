"""Experiment: segment size / compute lanes vs device-resident and end-to-end time (config 2 shape).
usage: python scripts/seg_exp.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pat_np, po_np, txt_np, to_np = workload.config2(n, seed=481)
pat = pkg.pinned_empty(len(pat_np), np.uint8); pat[:] = pat_np
txt = pkg.pinned_empty(len(txt_np), np.uint8); txt[:] = txt_np
po = pkg.pinned_empty(len(po_np), np.uint64); po[:] = po_np
to = pkg.pinned_empty(len(to_np), np.uint64); to[:] = to_np
res = pkg.pinned_empty(n, pkg.RESULT_DTYPE)
ref = {}
for lanes in (1, 2):
    for seg in (1 << 20, 1 << 18, 1 << 17, 1 << 16, 1 << 15):
        if seg > n and seg != 1 << 20:
            continue
        for mode in (0, 1):
            e = pkg.Engine(0)
            e.set_option(pkg.OPT_LANES, lanes)
            e.set_option(pkg.OPT_SEG_PAIRS, seg)
            e.set_option(pkg.OPT_SEG_BYTES, 1 << 40)
            e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
            e.run()
            best = None
            for _ in range(3):
                e.run()
                t = e.times()
                if best is None or t[2] < best[2]:
                    best = t
            r = e.download(n)
            if mode not in ref:
                ref[mode] = r.copy()
            ok = all(np.array_equal(r[f], ref[mode][f]) for f in r.dtype.names)
            e.align_packed(mode, pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
            w = []
            for _ in range(3):
                t0 = time.perf_counter()
                e.align_packed(mode, pat, po, txt, to, 1, -1, -1, want_ops=True, results=res)
                w.append((time.perf_counter() - t0) * 1e3)
            ok2 = all(np.array_equal(res[f], ref[mode][f]) for f in res.dtype.names)
            print(f"lanes {lanes} seg {seg:8d} mode {mode}: fill {best[0]:7.2f} tb {best[1]:7.2f} total {best[2]:7.2f} ms | e2e {min(w):7.2f} ms | same={ok and ok2}", flush=True)
            e.close()
