"""Generates tests/golden/hw3_kat.json from the UNMODIFIED reference binary oracle/_ref/hw3
(built by oracle/Makefile from /root/reference/Multiple_Sequence_Alignment/hw3.cpp).

hw3 prints no scores.  For TWO input sequences its PHYLIP output is exactly the optimal pairwise
alignment of its 3-state affine model (hw3.cpp:23-135), so the optimal score can be recomputed from
the alignment columns (tests/test_oracle.py:alignment_score) -- that pins the value the distance
stage (hw3.cpp:231-241) adds up.  For k > 2 sequences the first output row is the centre sequence,
which pins the arg-max of the sums.  Run here (the reference cannot travel):
    python tests/golden/make_golden_hw3.py
"""
import json
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HW3 = os.path.join(ROOT, "oracle", "_ref", "hw3")


def run_hw3(seqs, sc):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.fa"), os.path.join(td, "out.phy")
        with open(fin, "wb") as f:
            for i, s in enumerate(seqs):
                f.write(b">s%d\n" % i + s + b"\n")
        subprocess.check_call([HW3, "-i", fin, "-o", fout, "-s", "%d:%d:%d:%d" % sc], stdout=subprocess.DEVNULL)
        return open(fout).read()


def mutate(rng, base, alpha, k):
    s = bytearray(base)
    for _ in range(k):
        pos = rng.randrange(len(s))
        r = rng.random()
        if r < 0.5:
            s[pos] = rng.choice(alpha)
        elif r < 0.75 and len(s) > 3:
            del s[pos:pos + rng.randint(1, 4)]
        else:
            for _ in range(rng.randint(1, 4)):
                s.insert(pos, rng.choice(alpha))
    return bytes(s)


def main():
    rng = random.Random(4813)
    scorings = [(5, -4, -16, -4), (1, -1, -2, -1), (2, -3, -5, -2), (3, -1, 0, -2), (4, -6, -10, 0)]
    pairs, stars = [], []
    for it in range(160):
        alpha = rng.choice([b"ACGT", b"ACGT", b"AC", b"ACDEFGHIKLMNPQRSTVWY"])
        L = rng.choice([1, 2, 5, 20, 60, 150, 300])
        base = bytes(rng.choice(alpha) for _ in range(L))
        other = mutate(rng, base, alpha, rng.randint(0, max(1, L // 6)))
        if rng.random() < 0.15:
            other = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 2 * L)))
        sc = rng.choice(scorings)
        pairs.append({"seqs": [base.decode(), other.decode()], "s": list(sc), "phy": run_hw3([base, other], sc)})
    for it in range(30):
        k = rng.randint(3, 7)
        base = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(20, 120)))
        seqs = [mutate(rng, base, b"ACGT", rng.randint(0, 15)) for _ in range(k)]
        sc = rng.choice(scorings)
        stars.append({"seqs": [s.decode() for s in seqs], "s": list(sc), "phy": run_hw3(seqs, sc)})
    ship = os.path.join("/root/reference", "Multiple_Sequence_Alignment")
    shipped = {"fasta": open(os.path.join(ship, "input.fasta")).read(), "phy": open(os.path.join(ship, "output.phy")).read(), "s": [5, -4, -16, -4]}
    with tempfile.TemporaryDirectory() as td:      # the shipped golden must be what the binary writes today
        subprocess.check_call([HW3, "-i", os.path.join(ship, "input.fasta"), "-o", os.path.join(td, "o.phy"), "-s", "5:-4:-16:-4"], stdout=subprocess.DEVNULL)
        assert open(os.path.join(td, "o.phy")).read() == shipped["phy"]
    out = {"generator": "tests/golden/make_golden_hw3.py over oracle/_ref/hw3 (unmodified reference)", "pairs": pairs, "stars": stars, "shipped": shipped}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "hw3_kat.json"), "w"), indent=0)
    print("wrote", len(pairs), "pair vectors and", len(stars), "star vectors")


if __name__ == "__main__":
    sys.exit(main())
