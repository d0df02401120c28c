"""bin/hw2 end to end on config 2 written as FASTA (the drop-in CLI as a user runs it): wall time per mode.
usage: python scripts/cli_1m.py [pairs]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pat, po, txt, to = workload.config2(n, seed=481)
d = os.path.join(ROOT, "gpurun_out")
pf, tf = os.path.join(d, "cli_p.fa"), os.path.join(d, "cli_t.fa")
t0 = time.perf_counter()
workload.write_fasta(pf, pat, po, b"p"); workload.write_fasta(tf, txt, to, b"t")
print(f"wrote FASTA ({os.path.getsize(pf) + os.path.getsize(tf)} bytes) in {time.perf_counter() - t0:.1f} s", flush=True)
os.environ["HW2_TIMING"] = "1"
os.environ["B2A_TRACE"] = "0"
for flag, pin in (("-g", "0"), ("-g", "1"), ("-l", "0"), ("-l", "1")):
    os.environ["HW2_PIN"] = pin
    print("---- HW2_PIN =", pin, flush=True)
    for it in range(3):
        t0 = time.perf_counter()
        subprocess.check_call([pkg.HW2_BIN, flag, "-p", pf, "-t", tf, "-o", os.path.join(d, "cli_out.txt"), "-s", "1", "-1", "-1"])
        dt = time.perf_counter() - t0
        print(f"bin/hw2 {flag} {n} pairs: {dt:.2f} s wall = {n * 150 * 1000 / dt / 1e9:.0f} GCUPS (run {it})", flush=True)
    print(open(os.path.join(d, "cli_out.txt")).read().split("\n")[3:6])
for f in (pf, tf, os.path.join(d, "cli_out.txt")):
    os.remove(f)
