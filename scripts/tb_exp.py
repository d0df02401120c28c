import os, sys, json
sys.path.insert(0, ''+__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))+'')
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = 200000
pat, po, txt, to = workload.config2(n, seed=481)
for opt in (0, 1, 2, 3):
    os.environ["B2A_TB_OPT"] = str(opt)
    for mode in (0, 1):
        e = pkg.Engine(0)
        e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
        e.run()
        ts = [e.run() for _ in range(3)]
        print("opt", opt, "mode", mode, "fill %.2f tb %.2f" % (min(t[0] for t in ts), min(t[1] for t in ts)), flush=True)
        e.close()
