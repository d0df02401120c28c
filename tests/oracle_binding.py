"""ctypes binding to oracle/liboracle.so -- the CPU restatement of hw2.cpp.

TEST INFRASTRUCTURE ONLY (see oracle/hw2_oracle.c header). Builds the library
on first use with gcc if it is missing.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_HW2 = os.path.join(ORACLE_DIR, "_ref", "hw2")
REF_HW3 = os.path.join(ORACLE_DIR, "_ref", "hw3")
REF_HW4 = os.path.join(ORACLE_DIR, "_ref", "hw4")

GLOBAL, LOCAL = 0, 1


class OrcResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("end_i", C.c_uint32), ("end_j", C.c_uint32),
                ("start_i", C.c_uint32), ("start_j", C.c_uint32), ("overlap", C.c_int32),
                ("n_ops", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "hw2_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(so)
        _lib.orc_align.restype = C.c_int
        _lib.orc_align_ckpt.restype = C.c_int
        _lib.orc_score_only.restype = C.c_int
        _lib.orc_cigar.restype = C.c_size_t
        _lib.orc_mdz.restype = C.c_size_t
        _lib.orc_affine_score.restype = C.c_int
        _lib.orc_hw4_nw.restype = C.c_int
        _lib.orc_affine_align.restype = C.c_int
    return _lib


class Alignment:
    __slots__ = ("score", "end_i", "end_j", "start_i", "start_j", "overlap", "ops", "cigar", "mdz")

    def __repr__(self):
        return (f"Alignment(score={self.score}, end=({self.end_i},{self.end_j}), start=({self.start_i},{self.start_j}), "
                f"overlap={self.overlap}, cigar={self.cigar!r}, mdz={self.mdz!r})")


def align(mode, pattern: bytes, text: bytes, match: int, mismatch: int, gap: int) -> Alignment:
    """One pair through the oracle; returns score, coordinates, ops (traceback order), CIGAR, MD:Z."""
    L = lib()
    m, n = len(pattern), len(text)
    res = OrcResult()
    ops = C.create_string_buffer(m + n + 1)
    rc = L.orc_align(C.c_int(mode), pattern, C.c_uint32(m), text, C.c_uint32(n),
                     C.c_int(match), C.c_int(mismatch), C.c_int(gap), C.byref(res), ops)
    if rc != 0:
        raise RuntimeError(f"orc_align rc={rc}")
    a = Alignment()
    a.score, a.end_i, a.end_j = res.score, res.end_i, res.end_j
    a.start_i, a.start_j, a.overlap = res.start_i, res.start_j, res.overlap
    a.ops = ops.raw[:res.n_ops]
    cap = 16 * (m + n) + 64
    buf = C.create_string_buffer(cap)
    w = L.orc_cigar(a.ops, C.c_uint32(res.n_ops), buf, C.c_size_t(cap))
    a.cigar = buf.raw[:w].decode("latin-1")
    w = L.orc_mdz(a.ops, C.c_uint32(res.n_ops), pattern, text, C.c_uint32(res.start_i), C.c_uint32(res.start_j),
                  buf, C.c_size_t(cap))
    a.mdz = buf.raw[:w].decode("latin-1")
    return a


def align_ckpt(mode, pattern: bytes, text: bytes, match: int, mismatch: int, gap: int, ck: int = 1024) -> Alignment:
    """orc_align_ckpt: the same outputs as align() in O(ck * n + (m / ck) * n) memory (config 4 size)."""
    L = lib()
    m, n = len(pattern), len(text)
    res = OrcResult()
    ops = C.create_string_buffer(m + n + 1)
    rc = L.orc_align_ckpt(C.c_int(mode), pattern, C.c_uint32(m), text, C.c_uint32(n), C.c_int(match), C.c_int(mismatch), C.c_int(gap),
                          C.c_uint32(ck), C.byref(res), ops)
    if rc != 0:
        raise RuntimeError(f"orc_align_ckpt rc={rc}")
    a = Alignment()
    a.score, a.end_i, a.end_j = res.score, res.end_i, res.end_j
    a.start_i, a.start_j, a.overlap = res.start_i, res.start_j, res.overlap
    a.ops = ops.raw[:res.n_ops]
    cap = 16 * (m + n) + 64
    buf = C.create_string_buffer(cap)
    w = L.orc_cigar(a.ops, C.c_uint32(res.n_ops), buf, C.c_size_t(cap))
    a.cigar = buf.raw[:w].decode("latin-1")
    w = L.orc_mdz(a.ops, C.c_uint32(res.n_ops), pattern, text, C.c_uint32(res.start_i), C.c_uint32(res.start_j), buf, C.c_size_t(cap))
    a.mdz = buf.raw[:w].decode("latin-1")
    return a


def score_only(mode, pattern: bytes, text: bytes, match, mismatch, gap):
    L = lib()
    res = OrcResult()
    rc = L.orc_score_only(C.c_int(mode), pattern, C.c_uint32(len(pattern)), text, C.c_uint32(len(text)),
                          C.c_int(match), C.c_int(mismatch), C.c_int(gap), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orc_score_only rc={rc}")
    return res.score, res.end_i, res.end_j


def affine_score(s1: bytes, s2: bytes, match, mismatch, gopen, gext) -> int:
    L = lib()
    out = C.c_int32()
    rc = L.orc_affine_score(s1, C.c_uint32(len(s1)), s2, C.c_uint32(len(s2)), C.c_int(match), C.c_int(mismatch),
                            C.c_int(gopen), C.c_int(gext), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_affine_score rc={rc}")
    return out.value


def affine_align(s1: bytes, s2: bytes, match, mismatch, gopen, gext):
    """hw3's affine_alignment with traceback: (score, ops in traceback order)."""
    L = lib()
    score, nops = C.c_int32(), C.c_uint32()
    ops = C.create_string_buffer(len(s1) + len(s2) + 2)
    rc = L.orc_affine_align(s1, C.c_uint32(len(s1)), s2, C.c_uint32(len(s2)), C.c_int(match), C.c_int(mismatch),
                            C.c_int(gopen), C.c_int(gext), C.byref(score), C.byref(nops), ops)
    if rc != 0:
        raise RuntimeError(f"orc_affine_align rc={rc}")
    return score.value, ops.raw[:nops.value]


def aligned_rows(ops: bytes, s1: bytes, s2: bytes):
    """(alignmentString1, alignmentString2) of hw3.cpp:100-135 from an op list in traceback order"""
    a, b = bytearray(), bytearray()
    i = j = 0
    for op in reversed(ops):
        if op == 0x4D:
            a.append(s1[i]); b.append(s2[j]); i += 1; j += 1
        elif op == 0x44:
            a.append(s1[i]); b.append(0x2D); i += 1
        else:
            a.append(0x2D); b.append(s2[j]); j += 1
    return bytes(a), bytes(b)


def hw4_nw(s1: bytes, s2: bytes, match, mismatch, gap):
    """hw4's NW (tie order d > u > l): returns (score, distance, ops in traceback order)."""
    L = lib()
    score, dist, nops = C.c_int32(), C.c_int32(), C.c_uint32()
    ops = C.create_string_buffer(len(s1) + len(s2) + 1)
    rc = L.orc_hw4_nw(s1, C.c_uint32(len(s1)), s2, C.c_uint32(len(s2)), C.c_int(match), C.c_int(mismatch), C.c_int(gap),
                      C.byref(score), C.byref(dist), C.byref(nops), ops)
    if rc != 0:
        raise RuntimeError(f"orc_hw4_nw rc={rc}")
    return score.value, dist.value, ops.raw[:nops.value]


def select_best(mode, alignments):
    """hw2.cpp:326-357: strict '>' from -1000000, key = overlap (global) / score (local)."""
    best, idx = -1000000, -1
    for k, a in enumerate(alignments):
        key = a.overlap if mode == GLOBAL else a.score
        if key > best:
            best, idx = key, k
    return idx


def render_file(mode, patterns, texts, match, mismatch, gap) -> bytes:
    """The six-line output file hw2 writes (hw2.cpp:379-393), from oracle results."""
    als = [align(mode, p, t, match, mismatch, gap) for p, t in zip(patterns, texts)]
    k = select_best(mode, als)
    if k < 0:
        return b""
    head = b"Longest overlap:" if mode == GLOBAL else b"Highest local alignment score:"
    a = als[k]
    return b"\n".join([head, b"pattern=" + patterns[k], b"reference=" + texts[k],
                       b"Score =" + str(a.score).encode(), b"CIGAR =" + a.cigar.encode("latin-1"),
                       b"MD:Z=" + a.mdz.encode("latin-1")]) + b"\n"


def write_fasta(path, seqs, prefix):
    with open(path, "wb") as f:
        for k, s in enumerate(seqs):
            f.write(b">" + prefix + str(k).encode() + b"\n" + s + b"\n")


def run_hw2_binary(binary, flag, patterns, texts, match, mismatch, gap, tmpdir) -> bytes:
    """Run an hw2-compatible binary (the UNMODIFIED reference oracle/_ref/hw2, or the
    product's drop-in) on a batch; returns the bytes of the output file it wrote."""
    pf, tf, of = (os.path.join(str(tmpdir), x) for x in ("p.fa", "t.fa", "o.txt"))
    write_fasta(pf, patterns, b"p")
    write_fasta(tf, texts, b"t")
    if os.path.exists(of):
        os.remove(of)
    subprocess.check_call([binary, flag, "-p", pf, "-t", tf, "-o", of, "-s", str(match), str(mismatch), str(gap)])
    with open(of, "rb") as f:
        return f.read()


def have_ref():
    return os.path.exists(REF_HW2) and os.access(REF_HW2, os.X_OK)
