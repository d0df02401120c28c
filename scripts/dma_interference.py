"""Does a concurrent H2D copy slow the traceback kernel? resident run() with and without a side-stream H2D."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = 1000000
pat, po, txt, to = workload.config2(n, seed=481)
side = torch.cuda.Stream()
src = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
dst = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
back = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
for mode in (0, 1):
    e = pkg.Engine(0)
    e.set_option(pkg.OPT_SEG_PAIRS, 131072)
    e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
    e.run()
    for what in ("alone", "with H2D", "with D2H", "alone"):
        torch.cuda.synchronize()
        if what != "alone":
            with torch.cuda.stream(side):
                for _ in range(12):
                    if what == "with H2D":
                        dst.copy_(src, non_blocking=True)
                    else:
                        back.copy_(dst, non_blocking=True)
        e.run()
        print(mode, what, "fill %.2f tb %.2f total %.2f" % e.times(), flush=True)
        torch.cuda.synchronize()
    e.close()
