"""Seeded synthetic workloads of the shapes BASELINE.json names.

PRNG: numpy Generator(PCG64(seed)); the seed and this file are the record of the inputs.
  config 2: N pairs, text = 1000 iid uniform ACGT, pattern = 150-mer cut from the text at a uniform
            offset in [0, 850], 5 % substitution (uniform over ACGT, may be silent), 0.5 % deletion,
            0.5 % insertion, truncated / right-padded with iid bases to exactly 150.  seed 481.
  tie stress: a fraction of the pairs replaced by period-<=7 tandem repeats.
  config 4: one pair, text 100 kb iid, pattern = text with 8 % substitution, 1 % ins, 1 % del. seed 482.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def config2(n_pairs, seed=481, m=150, n=1000, psub=0.05, pdel=0.005, pins=0.005, tie_fraction=0.0, n_rate=0.0):
    """Returns (pat uint8[n_pairs*m], pat_off, txt uint8[n_pairs*n], txt_off) -- uniform shapes.
    n_rate > 0: that fraction of the PATTERN bases becomes 'N' afterwards (own generator: the other bases stay those of n_rate = 0)."""
    rng = np.random.default_rng(seed)
    txt = ACGT[rng.integers(0, 4, size=(n_pairs, n), dtype=np.uint8)]
    off = rng.integers(0, n - m + 1, size=n_pairs)
    src = txt[np.arange(n_pairs)[:, None], off[:, None] + np.arange(m)[None, :]]
    sub = rng.random((n_pairs, m)) < psub
    src = np.where(sub, ACGT[rng.integers(0, 4, size=(n_pairs, m), dtype=np.uint8)], src)
    dele = rng.random((n_pairs, m)) < pdel
    ins = rng.random((n_pairs, m)) < pins
    ins_base = ACGT[rng.integers(0, 4, size=(n_pairs, m), dtype=np.uint8)]
    pad = ACGT[rng.integers(0, 4, size=(n_pairs, m), dtype=np.uint8)]
    # output position of every kept source base / inserted base
    keep = ~dele
    width = ins.astype(np.int32) + keep.astype(np.int32)        # inserted base goes BEFORE the source base
    start = np.cumsum(width, axis=1) - width
    pat = pad.copy()
    rows = np.broadcast_to(np.arange(n_pairs)[:, None], (n_pairs, m))
    pos_ins = start
    ok = ins & (pos_ins < m)
    pat[rows[ok], pos_ins[ok]] = ins_base[ok]
    pos_keep = start + ins.astype(np.int32)
    ok = keep & (pos_keep < m)
    pat[rows[ok], pos_keep[ok]] = src[ok]
    # positions >= total length keep the iid padding; but padding must not survive INSIDE the edited prefix
    if tie_fraction > 0:
        k = int(n_pairs * tie_fraction)
        idx = rng.choice(n_pairs, size=k, replace=False)
        for i in idx:
            per = int(rng.integers(1, 8))
            unit = ACGT[rng.integers(0, 4, size=per, dtype=np.uint8)]
            txt[i] = np.tile(unit, n // per + 1)[:n]
            per2 = int(rng.integers(1, 8))
            unit2 = unit if rng.random() < 0.5 else ACGT[rng.integers(0, 4, size=per2, dtype=np.uint8)]
            pat[i] = np.tile(unit2, m // len(unit2) + 1)[:m]
    if n_rate > 0:
        pat[np.random.default_rng(seed + 7919).random(pat.shape) < n_rate] = ord("N")
    pat_off = (np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(m))
    txt_off = (np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(n))
    return np.ascontiguousarray(pat.reshape(-1)), pat_off, np.ascontiguousarray(txt.reshape(-1)), txt_off


def mutate(rng, seq, psub, pins, pdel):
    """seq uint8 -> mutated copy (per-base substitution / insertion-before / deletion)."""
    L = len(seq)
    s = np.where(rng.random(L) < psub, ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)], seq)
    keep = ~(rng.random(L) < pdel)
    ins = rng.random(L) < pins
    width = ins.astype(np.int64) + keep.astype(np.int64)
    start = np.cumsum(width) - width
    out = np.empty(int(width.sum()), dtype=np.uint8)
    out[start[ins]] = ACGT[rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)]
    out[(start + ins)[keep]] = s[keep]
    return out


def config4(length=100_000, seed=482):
    """One long pair: (pattern, text)."""
    rng = np.random.default_rng(seed)
    text = ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    return mutate(rng, text, 0.08, 0.01, 0.01), text


def config5(n_seqs=16, length=100_000, seed=483, divergence=0.10):
    """n_seqs diverged copies of a common ancestor (all-vs-all distance stage shape)."""
    rng = np.random.default_rng(seed)
    anc = ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    return [mutate(rng, anc, divergence * 0.8, divergence * 0.1, divergence * 0.1) for _ in range(n_seqs)]


def config5_shipped():
    """The reference's own 16 x 100 kb input (Multiple_Sequence_Alignment/input16100000.fasta, committed gzip'ed under
    tests/golden/): [(name, bytes)] read with hw3's rules (header = text after '>', all whitespace stripped, hw3.cpp:137-167)."""
    import gzip, os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input16100000.fasta.gz")
    out, name, cur = [], None, []
    for line in gzip.open(path, "rb").read().split(b"\n"):
        line = b"".join(line.split())
        if not line:
            continue
        if line.startswith(b">"):
            if name is not None:
                out.append((name, b"".join(cur)))
            name, cur = line[1:].decode("latin-1"), []
        else:
            cur.append(line)
    if name is not None:
        out.append((name, b"".join(cur)))
    return out


def split(data, off):
    """packed -> list of bytes"""
    return [data[int(off[k]):int(off[k + 1])].tobytes() for k in range(len(off) - 1)]


def write_fasta(path, data, off, prefix):
    with open(path, "wb") as f:
        for k in range(len(off) - 1):
            f.write(b">" + prefix + str(k).encode() + b"\n")
            f.write(data[int(off[k]):int(off[k + 1])].tobytes())
            f.write(b"\n")
