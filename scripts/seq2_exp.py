"""One b2a_align_batch_multi_seq2 call over config-2 pairs (for ncu -k regex:seq2_ : the expand / patch kernels' time and DRAM bytes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_rate = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
pat, po, txt, to = workload.config2(n, seed=481, n_rate=n_rate)
p2, t2 = pkg.PackedSeq(pat, pinned=True), pkg.PackedSeq(txt, pinned=True)
e = pkg.Engine(0)
modes = [pkg.GLOBAL, pkg.LOCAL]
for it in range(3):
    t0 = time.perf_counter()
    e.align_seq2_multi(modes, p2, po, t2, to, 1, -1, -1, want_ops=True)
    print("seq2 call %d: %.2f ms, h2d %d bytes, exceptions %d" % (it, (time.perf_counter() - t0) * 1e3, e.stats()["h2d_bytes"], p2.n_exc + t2.n_exc))
e.close()
