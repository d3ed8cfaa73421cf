#!/usr/bin/env python
"""Static SASS accounting for a kernel's hottest loop (no GPU needed).

usage: sass_loop.py <file.cubin|file.so> <function-name-substring> [--all]

Finds the innermost backward-branch loop with the most instructions (or the whole function with
--all), counts opcodes, and estimates issue cycles per warp on one SM sub-partition with the B200
rates measured by bench/pipe_peak.cu: IMAD.WIDE / IMAD.HI occupy the fma-heavy pipe for 4 cycles,
every other IMAD* for 2, ALU instructions (IADD3, LOP3, SHF, SEL, ISETP, MOV, ...) the alu pipe for 2.
"""
import re
import subprocess
import sys
from collections import Counter

HEAVY4 = ("IMAD.WIDE", "IMAD.HI")
ALU = ("IADD3", "LOP3", "SHF", "SEL", "ISETP", "MOV", "VIADD", "LEA", "PRMT", "IABS", "FSEL", "IADD", "VIMNMX", "PLOP3", "P2R", "R2P")


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


def account(instrs):
    cnt = Counter(opcode(t) for _, t in instrs)
    heavy = alu = other = 0
    for op, n in cnt.items():
        if op.startswith(HEAVY4):
            heavy += 4 * n
        elif op.startswith("IMAD"):
            heavy += 2 * n
        elif op.startswith(ALU):
            alu += 2 * n
        else:
            other += n
    return cnt, heavy, alu, other


def main():
    path, pat = sys.argv[1], sys.argv[2]
    whole = "--all" in sys.argv
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        body = ins
        if not whole:
            best = None
            for i, (addr, text) in enumerate(ins):
                m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*\s+.*0x([0-9a-f]+)", text)
                if opcode(text).startswith("BRA"):
                    m2 = re.search(r"0x([0-9a-f]+)", text)
                    if m2:
                        tgt = int(m2.group(1), 16)
                        if tgt <= addr:
                            seg = [x for x in ins if tgt <= x[0] <= addr]
                            if best is None or len(seg) > len(best):
                                best = seg
            if best:
                body = best
        cnt, heavy, alu, other = account(body)
        print(f"== {name}: {len(body)} instructions in {'function' if whole or body is ins else 'hottest loop'}")
        print("   " + ", ".join(f"{k}:{v}" for k, v in cnt.most_common()))
        print(f"   est. cycles/warp/SMSP: fma-heavy {heavy}, alu {alu}, other-issue {other}, issue slots {len(body)}")


if __name__ == "__main__":
    main()
