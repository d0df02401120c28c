"""Traceback kernel: L2-prefetch option sweep (B2A_OPT_TB bits 1 / 2 / 4, csrc/traceback.cuh) on the config-2 shape.
usage: python scripts/tb_exp.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
pat, po, txt, to = workload.config2(n, seed=481)
for mode in (0, 1):
    e = pkg.Engine(0)
    e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
    e.run()
    ref = None
    for opt in (0, 1, 2, 3, 4, 6, 7):
        e.set_option(pkg.OPT_TB, opt)
        e.run()
        ts = [e.run() for _ in range(3)]
        res = e.download(n)
        key = res.tobytes()
        ref = ref or key
        print("mode", mode, "opt", opt, "fill %.3f tb %.3f ms" % (min(t[0] for t in ts), min(t[1] for t in ts)), "same" if key == ref else "DIFFERENT", flush=True)
    e.close()
