"""FASTA-dialect vectors for the hw2 drop-in: raw pattern/text FILE CONTENTS (CRLF, blank lines, records with empty
bodies, text before the first header, trailing whitespace, lower case, no final newline ...) and what the UNMODIFIED
reference binary (oracle/_ref/hw2) does with them: exit code, stderr, output bytes.  -> tests/golden/hw2_fasta_kat.json
    python tests/golden/make_golden_fasta.py
"""
import json, os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HW2 = os.path.join(ROOT, "oracle", "_ref", "hw2")

FILES = [
    # (patterns file, texts file)
    (">p0\r\nACGTACGT\r\n>p1\r\nTTGACC\r\n", ">t0\r\nACGTTCGT\r\n>t1\r\nTTGGACC\r\n"),                      # CRLF
    (">p0\nACGT\nACGT\n\n>p1\n\nTTG\nACC", ">t0\nACGTT\nCGT\n>t1\nTTGGACC"),                                  # wrapped lines, blank lines, no final newline
    ("ACGTAC\n>p1\nGGGTTT\n", ">t0\nACGAAC\n>t1\nGGCTTT\n"),                                               # text before the first header is a record
    (">p0\n>p1\nACGT\n>p2\nTTTT\n", ">t0\nACGA\n>t1\n\n>t2\nTTAT\n"),                                       # empty records vanish (and shift the pairing)
    (">p0\nACGT  \t\n>p1\n  AC GT\n", ">t0\nACGT\n>t1\n  AC GT\n"),                                          # trailing blanks stripped, inner/leading kept
    (">p0\nacgt\n>p1\nACGT\n", ">t0\nACGT\n>t1\nACGT\n"),                                                   # case-sensitive comparison
    (">p0\nACGT\n", ">t0\nACGT\n>t1\nGGGG\n"),                                                             # count mismatch -> error, no output file
    ("", ""),                                                                                             # no records: empty output file
    (">only header\n", ">only header\n"),
    (">p0\nAC-GT\n>p1\nNNNN\n", ">t0\nAC-GT\n>t1\nNNAN\n"),                                                # '-' and N are ordinary bytes (overlap treats '-' specially)
]


def main():
    out = []
    for pf, tf in FILES:
        for flag in ("-g", "-l"):
            for sc in ((1, -1, -1), (2, -3, -4)):
                with tempfile.TemporaryDirectory() as td:
                    open(os.path.join(td, "p.fa"), "w", newline="").write(pf)
                    open(os.path.join(td, "t.fa"), "w", newline="").write(tf)
                    p = subprocess.run([HW2, flag, "-p", "p.fa", "-t", "t.fa", "-o", "o.txt", "-s", *map(str, sc)], cwd=td, capture_output=True, text=True)
                    body = open(os.path.join(td, "o.txt")).read() if os.path.exists(os.path.join(td, "o.txt")) else None
                out.append({"patterns": pf, "texts": tf, "flag": flag, "s": list(sc), "rc": p.returncode, "stderr": p.stderr, "output": body})
    json.dump({"generator": "tests/golden/make_golden_fasta.py over oracle/_ref/hw2 (unmodified reference)", "cases": out},
              open(os.path.join(ROOT, "tests", "golden", "hw2_fasta_kat.json"), "w"), indent=0)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    sys.exit(main())
