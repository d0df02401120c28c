"""Traceback kernel occupancy variants (TB_MIN_CTAS builds under bioinformatics-algorithms_b200/variants/): fill / traceback ms per mode.
usage: python scripts/tb_variants.py [pairs]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "400000"
child = r'''
import os, sys
sys.path.insert(0, %r)
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
n = int(sys.argv[1])
pat, po, txt, to = workload.config2(n, seed=481)
for mode in (0, 1):
    e = pkg.Engine(0)
    e.upload(mode, pat, po, txt, to, 1, -1, -1, want_ops=True)
    e.run()
    best = min(((e.run(), e.times())[1] for _ in range(3)), key=lambda t: t[1])
    res = e.download(n)
    print("  mode", mode, "fill %%.3f tb %%.3f total %%.3f ms; checksum" %% best, int(res["score"].astype("i8").sum()), int(res["n_ops"].astype("i8").sum()), int(res["overlap"].astype("i8").sum()), flush=True)
    e.close()
''' % ROOT
libs = {"default": None}
vdir = os.path.join(ROOT, "bioinformatics-algorithms_b200", "variants")
if os.path.isdir(vdir):
    for f in sorted(os.listdir(vdir)):
        if f.endswith(".so"):
            libs[f] = os.path.join(vdir, f)
for name, path in libs.items():
    env = dict(os.environ)
    if path:
        env["B2A_LIB"] = path
    print(name, flush=True)
    subprocess.run([sys.executable, "-c", child, n], env=env)
