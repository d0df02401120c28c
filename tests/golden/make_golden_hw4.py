"""Generates tests/golden/hw4_kat.json from the UNMODIFIED reference binary oracle/_ref/hw4
(built by oracle/Makefile from /root/reference/hw4/hw4.cpp).

hw4 prints only the Newick tree.  For TWO sequences the tree is "(a:h,b:h):0.0;" with h = distance / 2
(hw4.cpp:179-190), which pins the NW distance of the pair (tie order d > u > l, hw4.cpp:37-46); for more sequences
the whole string pins the distance matrix + UPGMA + number formatting.  Includes the shipped input.fasta/tree.txt.
    python tests/golden/make_golden_hw4.py
"""
import json, os, random, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HW4 = os.path.join(ROOT, "oracle", "_ref", "hw4")


def run_hw4(fasta_text, sc):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.fa"), os.path.join(td, "tree.txt")
        open(fin, "w").write(fasta_text)
        subprocess.check_call([HW4, "-i", fin, "-t", fout, "-s", *map(str, sc)])
        return open(fout).read()


def fasta(seqs, names=None):
    return "".join(">%s\n%s\n" % (names[i] if names else "s%d" % i, s) for i, s in enumerate(seqs))


def mutate(rng, base, alpha, k):
    s = list(base)
    for _ in range(k):
        pos = rng.randrange(len(s))
        r = rng.random()
        if r < 0.5:
            s[pos] = rng.choice(alpha)
        elif r < 0.75 and len(s) > 3:
            del s[pos:pos + rng.randint(1, 3)]
        else:
            for _ in range(rng.randint(1, 3)):
                s.insert(pos, rng.choice(alpha))
    return "".join(s)


def main():
    rng = random.Random(4814)
    scorings = [(1, -1, -1), (2, -3, -4), (5, -4, -16), (1, -1, -2), (3, -2, -1), (1, 0, 0)]
    pairs, trees = [], []
    for it in range(200):
        alpha = rng.choice(["ACGT", "ACGT", "AC", "A", "ACDEFGHIKLMNPQRSTVWY"])
        L = rng.choice([1, 2, 3, 5, 8, 20, 60, 150, 300])
        a = "".join(rng.choice(alpha) for _ in range(L))
        b = mutate(rng, a, alpha, rng.randint(0, max(1, L // 5)))
        if rng.random() < 0.2:
            b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 2 * L)))
        if rng.random() < 0.15:                      # tie stress: tandem repeats
            u = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 4)))
            a, b = (u * 40)[:rng.randint(5, 80)], (u * 40)[:rng.randint(5, 80)]
        sc = rng.choice(scorings)
        pairs.append({"seqs": [a, b], "s": list(sc), "tree": run_hw4(fasta([a, b]), sc)})
    for it in range(40):
        k = rng.randint(3, 9)
        base = "".join(rng.choice("ACGT") for _ in range(rng.randint(5, 120)))
        seqs = [mutate(rng, base, "ACGT", rng.randint(0, 12)) for _ in range(k)]
        if rng.random() < 0.3:
            seqs[rng.randrange(k)] = seqs[0]          # identical sequences -> zero distances, UPGMA ties
        sc = rng.choice(scorings)
        trees.append({"seqs": seqs, "s": list(sc), "tree": run_hw4(fasta(seqs), sc)})
    ship_dir = "/root/reference/hw4"
    shipped = {"fasta": open(os.path.join(ship_dir, "input.fasta")).read(), "tree": open(os.path.join(ship_dir, "tree.txt")).read(), "s": [1, -1, -1]}
    assert run_hw4(shipped["fasta"], (1, -1, -1)) == shipped["tree"]
    out = {"generator": "tests/golden/make_golden_hw4.py over oracle/_ref/hw4 (unmodified reference)", "pairs": pairs, "trees": trees, "shipped": shipped}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "hw4_kat.json"), "w"), indent=0)
    print("wrote", len(pairs), "pair vectors and", len(trees), "tree vectors")


if __name__ == "__main__":
    sys.exit(main())
