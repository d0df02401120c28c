"""Config 4 golden: the UNMODIFIED reference binary (oracle/_ref/hw2, built from /root/reference by oracle/Makefile) on the
seed-482 100 kb x 100 kb pair, -l and -g, -s 1 -1 -1.  Needs ~50 GB of RAM and a few minutes per mode (hw2.cpp:119-120
allocates the full int + char matrices).  Writes tests/golden/c4_local.txt.gz / c4_global.txt.gz (the 6-line output files).

    python tests/golden/make_golden_c4.py [workdir]
"""
import gzip, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
load_package()
from bioinformatics_algorithms_b200 import workload

work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/c4"
os.makedirs(work, exist_ok=True)
p, t = workload.config4(100_000, seed=482)
for name, seq in (("p.fa", p), ("t.fa", t)):
    with open(os.path.join(work, name), "wb") as f:
        f.write(b">s\n" + seq.tobytes() + b"\n")
ref = os.path.join(ROOT, "oracle", "_ref", "hw2")
for flag, out in (("-l", "c4_local.txt"), ("-g", "c4_global.txt")):
    t0 = time.time()
    subprocess.check_call([ref, flag, "-p", os.path.join(work, "p.fa"), "-t", os.path.join(work, "t.fa"),
                           "-o", os.path.join(work, out), "-s", "1", "-1", "-1"])
    print(flag, "reference hw2 took %.1f s" % (time.time() - t0), flush=True)
    with open(os.path.join(work, out), "rb") as f, gzip.open(os.path.join(ROOT, "tests", "golden", out + ".gz"), "wb", 9) as g:
        g.write(f.read())
