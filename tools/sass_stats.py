#!/usr/bin/env python
"""Instruction-mix report from `cuobjdump -sass`: per kernel, per loop (backward branch).

Used to replace the op-count estimates of SURVEY.md §8(d) with SASS counts, and to see
which pipe (FMA / ALU / MUFU / LSU) bounds the light loop before spending GPU time.

    python tools/sass_stats.py svbrdf_diff_renderer_b200/csrc/libsvbrdf_b200.so [name-filter]
"""
import collections
import re
import subprocess
import sys

FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HMUL2", "HADD2", "DFMA", "DMUL", "DADD"}
ALU = {"FMNMX", "FSEL", "FSETP", "FSET", "LOP3", "IADD3", "IADD", "SHF", "PRMT", "ISETP", "SEL", "MOV", "LEA", "I2F", "F2I",
       "I2FP", "F2FP", "FCHK", "PLOP3", "VIADD", "IABS", "P2R", "R2P", "FMNMX3", "UMOV", "I2F.U16", "CS2R", "S2R", "FSWZADD"}
LSU = {"LDG", "STG", "LDS", "STS", "LDC", "LDCU", "ULDC", "LD", "ST", "LDL", "STL", "ATOMG", "RED", "LDSM", "SHFL", "UBLKCP", "SYNCS"}


def classify(op):
    base = op.split(".")[0]
    if base == "MUFU":
        return "MUFU"
    if base in FMA:
        return "FMA"
    if base in LSU:
        return "LSU"
    if base in ("BRA", "EXIT", "BAR", "BSSY", "BSYNC", "CALL", "RET", "NOP", "WARPSYNC", "BMOV", "BREAK", "YIELD", "DEPBAR", "ERRBAR", "MEMBAR"):
        return "CTRL"
    return "ALU"


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if flt and flt not in name:
            continue
        ins = []  # (addr, op, text)
        for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)\s*([^;]*);", f):
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
        total = collections.Counter(classify(op) for _, op, _ in ins)
        print(f"== {name}\n   total {len(ins)} instr: {dict(total)}")
        # loops = backward branches
        loops = []
        for addr, op, rest in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) <= addr:
                    loops.append((int(t.group(1), 16), addr))
        for lo, hi in loops:
            body = [(a, op) for a, op, _ in ins if lo <= a <= hi]
            cnt = collections.Counter(classify(op) for _, op in body)
            mufu = collections.Counter(op for _, op in body if op.startswith("MUFU"))
            mem = collections.Counter(op.split(".")[0] for _, op in body if classify(op) == "LSU")
            print(f"   loop 0x{lo:04x}-0x{hi:04x}: {len(body):4d} instr  {dict(cnt)}  mufu={dict(mufu)} mem={dict(mem)}")


if __name__ == "__main__":
    main()
