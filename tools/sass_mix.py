#!/usr/bin/env python3
"""Static pipe mix of the synthesis kernel's unrolled sample loop.

    python tools/sass_mix.py [lib.so] [--dump]

Finds e1_synth_kernel<16> in the library's SASS, takes the straight-line region from the first
carrier-index multiply (IMAD.WIDE.U32 ..., 0x1ff, ...) to the last accumulator IMAD of the run and
counts warp instructions per issue pipe (Blackwell: IMAD* = fmaheavy, 1 per 2 clk per SMSP; IADD3 /
LOP3 / SHF / ISETP / VIMNMX / VIADD / LEA / SEL / PRMT = alu, 1 per 2 clk; LDS = lsu).  Used here,
without a GPU, to compare formulations before spending GPU time on them."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEAVY = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP")
ALU = ("IADD3", "LOP3", "SHF", "ISETP", "VIMNMX", "VIMNMX3", "VIADD", "VIADDMNMX", "LEA", "SEL", "PRMT", "PLOP3", "MOV", "IABS", "SGXT", "BMSK",
       "FMNMX", "FSETP", "FSEL", "I2FP", "F2FP", "VABSDIFF", "IMNMX")


def pipe(op):
    base = op.split(".")[0]
    if base in HEAVY:
        return "heavy"
    if base in ALU:
        return "alu"
    if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "ATOMG", "RED"):
        return "lsu"
    if base in ("BRA", "BSSY", "BSYNC", "BREAK", "CALL", "RET", "EXIT", "WARPSYNC", "BAR"):
        return "cbu"
    if base.startswith("U") or base in ("S2UR", "R2UR", "LDCU"):
        return "uniform"
    return "other:" + base


def function_sass(lib, name="e1_synth_kernelILi16E"):
    txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = name in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = Path(args[0]) if args else ROOT / "galileo-sdr-sim_b200" / "lib" / "libe1b200.so"
    ins = function_sass(lib)
    # straight-line regions (cut at branches); the unrolled sample loops are the ones with >= 8 LDS
    regions, cur = [], []
    for a, t in ins:
        toks = t.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        cur.append((a, t))
        if op.startswith(("BRA", "BSYNC", "CALL", "RET", "EXIT")):
            regions.append(cur)
            cur = []
    regions.append(cur)
    n = 16.0
    for region in regions:
        if sum(1 for _, t in region if " LDS" in " " + t) < 8:
            continue
        c, ops = Counter(), Counter()
        for _, t in region:
            toks = t.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            c[pipe(op)] += 1
            ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD.") else "")] += 1
        print(f"{lib.name}: region {region[0][0]:#x}..{region[-1][0]:#x}, {len(region)} instr, per sample: "
              + ", ".join(f"{k} {v / n:.2f}" for k, v in sorted(c.items())))
        print("   clk/sample-warp >= max(2*heavy, 2*alu, total) = %.1f" % max(2 * c["heavy"] / n, 2 * c["alu"] / n, len(region) / n))
        print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
        if "--dump" in sys.argv:
            for a, t in region:
                print(f"{a:05x}  {t}")
    print(f"   whole kernel: {len(ins)} instr")
    if "--dump" in sys.argv:
        for a, t in region:
            print(f"{a:05x}  {t}")


if __name__ == "__main__":
    main()
