#!/usr/bin/env python
"""Innermost-loop opcode census of a kernel from `cuobjdump -sass` output.

    cuobjdump -sass build/qw_horner.o | python tools/sass_loops.py 'rk4_cycle_hornerILi16ELi64ELi6ELb0ELb0'

For every backward branch of the selected function prints the loop body's size and its FP64 /
memory / barrier opcode counts (the step loops of the RK4 kernels are the bodies with four
BAR.SYNC, or none in the barrier-free kernels).  This is how the "executed FP64 instructions
per site-update" figures in DESIGN.md are checked without a GPU; the ncu opcode counters in
profiles/ are the run-time counterpart.
"""
import re
import sys
from collections import Counter


def functions(lines):
    name, body = None, []
    for ln in lines:
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            body.append(ln)
    if name:
        yield name, body


def main():
    pat = sys.argv[1]
    text = sys.stdin.read().splitlines()
    for name, body in functions(text):
        if pat not in name:
            continue
        ins = []
        for ln in body:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print("== %s: %d instructions" % (name, len(ins)))
        tot = Counter(op_of(t) for _, t in ins)
        print("   whole kernel:", fmt(tot))
        for addr, txt in ins:
            m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", txt)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                loop = [t for a, t in ins if tgt <= a <= addr]
                c = Counter(op_of(t) for t in loop)
                print("   loop %#06x..%#06x: %5d instr  %s" % (tgt, addr, len(loop), fmt(c)))


def op_of(txt):
    t = re.sub(r"^@!?U?P\d+\s+", "", txt)
    return t.split()[0].split(".")[0]


KEYS = ["DFMA", "DADD", "DMUL", "DSETP", "SHFL", "LDS", "STS", "LDL", "STL", "LDG", "STG", "BAR",
        "SEL", "FSEL", "MOV", "IMAD", "UBLKCP", "SYNCS", "LDGSTS"]


def fmt(c):
    return " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])


if __name__ == "__main__":
    main()
