#!/usr/bin/env python3
"""Regenerates tests/golden/hw2_kat.json from the UNMODIFIED reference binary.

Run in the build container (needs oracle/_ref/hw2, which oracle/Makefile compiles
from /root/reference/Local_Global_Alignment/hw2.cpp where it lies):

    make -C oracle && python tests/golden/make_golden.py

Each case is a 1-pair batch (hw2 prints only the batch winner, hw2.cpp:379-393),
so the Score/CIGAR/MD:Z lines are that pair's own result.  The shipped fixtures
(patterns.fasta x texts.fasta -> global.txt / local.txt) are recorded verbatim
as multi-pair cases.  The GPU box has no /root/reference: tests read this JSON.
"""
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as ob  # noqa: E402

REF = "/root/reference/Local_Global_Alignment"


def read_fasta_simple(path):
    seqs, cur = [], b""
    for line in open(path, "rb").read().split(b"\n"):
        line = line.rstrip()
        if not line:
            continue
        if line.startswith(b">"):
            if cur:
                seqs.append(cur)
                cur = b""
        else:
            cur += line
    if cur:
        seqs.append(cur)
    return seqs


def parse(out: bytes):
    lines = out.decode("latin-1").split("\n")
    if len(lines) < 6:
        return None
    return {"header": lines[0], "pattern": lines[1][len("pattern="):], "reference": lines[2][len("reference="):],
            "score": int(lines[3][len("Score ="):]), "cigar": lines[4][len("CIGAR ="):], "mdz": lines[5][len("MD:Z="):]}


def rand_seq(rng, n, alphabet):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, alphabet, psub, pindel):
    out = []
    for ch in s:
        r = rng.random()
        if r < pindel:
            continue
        if r < 2 * pindel:
            out.append(rng.choice(alphabet))
        out.append(rng.choice(alphabet) if rng.random() < psub else ch)
    return "".join(out)


def main():
    if not ob.have_ref():
        sys.exit("oracle/_ref/hw2 missing: run `make -C oracle` in the build container first")
    rng = random.Random(481)
    single = []
    # SURVEY.md Appendix B vectors (inputs only; outputs come from the binary below)
    appendix_b = [
        ("g", "A", "AA", (1, -1, -1)), ("g", "AA", "A", (1, -1, -1)), ("g", "AB", "BA", (1, -1, -1)),
        ("g", "ACGT", "TGCA", (1, -1, -1)), ("g", "AAAA", "AA", (1, -1, -1)), ("g", "AA", "AAAA", (1, -1, -1)),
        ("g", "ACAC", "CACA", (1, -1, -1)), ("g", "AAAA", "CCCC", (1, -1, -1)), ("g", "GATTACA", "GCATGCT", (1, -1, -2)),
        ("g", "ACGTTTACGT", "ACGTACGT", (3, -2, -1)), ("l", "AB", "BA", (1, -1, -1)), ("l", "ACAC", "CACA", (1, -1, -1)),
        ("l", "AATAA", "AAGAA", (2, -1, -1)), ("l", "AAGAA", "AAAA", (2, -1, -1)), ("l", "AAAA", "AAGAA", (2, -1, -1)),
        ("l", "ACGTACGT", "ACGT", (1, -1, -1)), ("l", "acgt", "ACGT", (1, -1, -1)), ("l", "AAAA", "CCCC", (1, -1, -1)),
        ("l", "ACGTTTACGT", "ACGTACGT", (3, -2, -1)), ("l", "GATTACA", "GCATGCT", (1, -1, -2)),
    ]
    cases = list(appendix_b)
    scorings = [(1, -1, -1), (2, -3, -4), (5, -4, -16), (3, -2, -1), (1, -1, -2), (2, -1, -1), (1, 0, 0), (4, -6, -3)]
    for _ in range(420):
        kind = rng.random()
        if kind < 0.25:      # tie stress: tiny alphabets / homopolymers / tandem repeats
            unit = rand_seq(rng, rng.randint(1, 4), "AC")
            p = (unit * 40)[: rng.randint(1, 36)]
            t = (rand_seq(rng, rng.randint(1, 4), "AC") * 40)[: rng.randint(1, 48)]
        elif kind < 0.6:     # related sequences (pattern = mutated slice of text)
            t = rand_seq(rng, rng.randint(8, 70), "ACGT")
            a = rng.randint(0, max(0, len(t) - 4))
            p = mutate(rng, t[a: a + rng.randint(3, 40)], "ACGT", 0.1, 0.05) or "A"
        elif kind < 0.8:     # unrelated
            p = rand_seq(rng, rng.randint(1, 40), "ACGT")
            t = rand_seq(rng, rng.randint(1, 64), "ACGT")
        else:                # general byte alphabet (protein-like, case-sensitive)
            p = rand_seq(rng, rng.randint(1, 30), "ACDEFGHIKLMNacgt")
            t = rand_seq(rng, rng.randint(1, 40), "ACDEFGHIKLMNacgt")
        cases.append((rng.choice("gl"), p, t, rng.choice(scorings)))
    with tempfile.TemporaryDirectory() as td:
        for mode, p, t, s in cases:
            out = ob.run_hw2_binary(ob.REF_HW2, "-" + mode, [p.encode()], [t.encode()], *s, td)
            r = parse(out)
            assert r is not None and r["pattern"] == p and r["reference"] == t
            single.append({"mode": mode, "p": p, "t": t, "s": list(s),
                           "score": r["score"], "cigar": r["cigar"], "mdz": r["mdz"]})
        # multi-pair cases exercise the selection rule (hw2.cpp:340-357)
        multi = []
        for _ in range(40):
            k = rng.randint(2, 6)
            ps, ts = [], []
            for _ in range(k):
                t = rand_seq(rng, rng.randint(4, 40), "ACGT")
                ts.append(t)
                ps.append(mutate(rng, t[: rng.randint(2, 30)], "ACGT", 0.15, 0.05) or "C")
            mode, s = rng.choice("gl"), rng.choice(scorings[:6])
            out = ob.run_hw2_binary(ob.REF_HW2, "-" + mode, [x.encode() for x in ps], [x.encode() for x in ts], *s, td)
            multi.append({"mode": mode, "patterns": ps, "texts": ts, "s": list(s), "output": out.decode("latin-1")})
    shipped = {
        "patterns": [x.decode() for x in read_fasta_simple(os.path.join(REF, "patterns.fasta"))],
        "texts": [x.decode() for x in read_fasta_simple(os.path.join(REF, "texts.fasta"))],
        "patterns_fasta": open(os.path.join(REF, "patterns.fasta"), "rb").read().decode("latin-1"),
        "texts_fasta": open(os.path.join(REF, "texts.fasta"), "rb").read().decode("latin-1"),
        "s": [1, -1, -1],
        "global_txt": open(os.path.join(REF, "global.txt"), "rb").read().decode("latin-1"),
        "local_txt": open(os.path.join(REF, "local.txt"), "rb").read().decode("latin-1"),
    }
    with open(os.path.join(HERE, "hw2_kat.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (seed 481) via oracle/_ref/hw2",
                   "single": single, "multi": multi, "shipped": shipped}, f, indent=0)
    print(f"wrote {len(single)} single-pair, {len(multi)} multi-pair cases + shipped fixture")


if __name__ == "__main__":
    main()
