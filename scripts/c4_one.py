"""One config-4 launch (seed-482 100 kb x 100 kb pair, local, score + traceback) for ncu: python scripts/c4_one.py [len] [runs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p, t = workload.config4(L, seed=482)
pat, po = pkg.pack([p.tobytes()]); txt, to = pkg.pack([t.tobytes()])
e = pkg.Engine(0)
e.upload(pkg.LOCAL, pat, po, txt, to, 1, -1, -1, want_ops=True)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    e.run(); print(e.times())
e.close()
