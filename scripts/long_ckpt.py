"""One very long pair through checkpointed recomputation (score pass + band groups re-filled bottom-up): time, memory, and the op list
re-scored on the host.  usage: python scripts/long_ckpt.py [length=1000000] [mode=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p, t = workload.config4(L, seed=482)
pat, po = pkg.pack([p.tobytes()]); txt, to = pkg.pack([t.tobytes()])
e = pkg.Engine(0)
import torch
free0 = torch.cuda.mem_get_info()[0]
t0 = time.perf_counter()
res = e.align_packed(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
dt = time.perf_counter() - t0
free1 = torch.cuda.mem_get_info()[0]
words, off = e.copy_ops(1)
ops = np.frombuffer(pkg.unpack_ops(words, off, 0, res["n_ops"][0]), dtype=np.uint8)
di = (ops != 0x49).astype(np.int64); dj = (ops != 0x44).astype(np.int64)
i = int(res["end_i"][0]) - np.cumsum(di); j = int(res["end_j"][0]) - np.cumsum(dj)
isM = ops == 0x4D
eq = p[i[isM]] == t[j[isM]]
score = int(eq.sum()) - int((~eq).sum()) - int((~isM).sum())
print(f"{len(p)} x {len(t)} mode {mode}: {dt:.2f} s wall = {len(p) * len(t) / dt / 1e9:.0f} GCUPS incl. both passes; device memory held {(free0 - free1) / 2**30:.1f} GiB "
      f"(a stored record would need {len(p) * len(t) * 0.5 / 2**30:.0f} GiB); score {int(res['score'][0])}, n_ops {int(res['n_ops'][0])}, "
      f"ops re-scored {score} {'OK' if score == int(res['score'][0]) and (int(i[-1]), int(j[-1])) == (int(res['start_i'][0]), int(res['start_j'][0])) else 'MISMATCH'}, "
      f"stats {e.stats()}")
e.close()
