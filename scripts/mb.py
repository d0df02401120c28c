import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
e = pkg.Engine(0)
names = {0: "VIADDMNMX.S16x2", 1: "fill ALU mix", 2: "mix + IMAD", 3: "VIMNMX.S16x2 (plain max)", 4: "VIMNMX3.S16x2", 5: "VIADD.16x2",
         6: "PRMT", 7: "LOP3", 8: "IADD", 9: "IMAD", 10: "NW cell pair (3 ALU + IMAD)", 11: "SW cell pair (5 ALU + IMAD)"}
for k in range(12):
    g, mhz = e.microbench(k)
    print(f"kind {k} {names[k]:28s}: {g:9.0f} G lane-instr/s = {g * 1e9 / (148 * mhz * 1e6):6.1f} lanes/clk/SM at {mhz:.0f} MHz")
e.close()
