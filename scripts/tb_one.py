import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pat, po, txt, to = workload.config2(n, seed=481)
e = pkg.Engine(0)
e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
for _ in range(2):
    print(e.run())
e.close()
