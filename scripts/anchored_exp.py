"""Config-4-shaped pair (100 kb x 100 kb, 8 % substitutions, 1 % indels), GLOBAL: the full alignment against the seed-anchored one."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload
length = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pat, txt = workload.config4(length, seed=482)
pat, txt = bytes(pat), bytes(txt)
e = pkg.Engine(0)
def rescore(ops, s=(1, -1, -1)):
    i = j = sc = 0
    P, T = np.frombuffer(pat, np.uint8), np.frombuffer(txt, np.uint8)
    o = np.frombuffer(ops, np.uint8)[::-1]
    isM, isD = o == 0x4D, o == 0x44
    pi = np.cumsum(isM | isD) - 1; tj = np.cumsum(isM | ~isD) - 1
    eq = P[pi[isM]] == T[tj[isM]]
    assert pi[-1] == len(P) - 1 and tj[-1] == len(T) - 1
    return int(eq.sum() * s[0] + (~eq).sum() * s[1] + (~isM).sum() * s[2])
for it in range(3):
    t0 = time.perf_counter()
    res, ops = e.align_batch(pkg.GLOBAL, [pat], [txt], 1, -1, -1, want_ops=True)
    t_full = time.perf_counter() - t0
print("full    : score %d, n_ops %d, %.1f ms (host call incl. op download)" % (res["score"][0], res["n_ops"][0], t_full * 1e3))
assert rescore(ops[0]) == int(res["score"][0])
for k, spacing in ((16, 256), (16, 512), (16, 2048), (20, 8192)):
    t0 = time.perf_counter()
    a = pkg.find_anchors(pat, txt, k, spacing)
    t_find = time.perf_counter() - t0
    for it in range(3):
        t0 = time.perf_counter()
        r, aops = e.align_anchored(pat, txt, a, 1, -1, -1)
        t_al = time.perf_counter() - t0
    assert rescore(aops) == int(r["score"])
    cells = sum(int(x) * int(y) for x, y in zip(np.diff(np.concatenate([[0], a["i"] + a["len"], [len(pat)]])) , np.diff(np.concatenate([[0], a["j"] + a["len"], [len(txt)]]))))
    print("anchored: k %2d spacing %5d: %4d anchors in %.1f ms, score %d (full %d), n_ops %d, align %.2f ms, ~%.2e of %.2e cells" %
          (k, spacing, len(a), t_find * 1e3, r["score"], res["score"][0], r["n_ops"], t_al * 1e3, cells, len(pat) * len(txt)))
e.close()
