#!/usr/bin/env python3
"""Opcode mix per pipe for kernels matching a filter: tools/sass_mix.py <file.so> <filter>"""
import collections, re, subprocess, sys
out = subprocess.check_output(["cuobjdump", "-sass", sys.argv[1]], text=True)
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ALU = ("IADD3", "LOP3", "SHF", "PRMT", "SEL", "ISETP", "VIADD", "MOV", "PLOP3", "LEA", "IABS", "FMNMX", "VIMNMX")
FMA = ("IMAD", "FFMA", "HFMA2", "FMUL", "FADD")
f, cnt = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        f = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and f:
        op = m.group(2)
        base = op.split(".")[0]
        key = base + (".WIDE" if ".WIDE" in op else "")
        cnt[f][key] += 1
for f, c in cnt.items():
    if flt in f:
        alu = sum(v for k, v in c.items() if k.split(".")[0] in ALU)
        fma = sum(v for k, v in c.items() if k.split(".")[0] in FMA and ".WIDE" not in k)
        wide = sum(v for k, v in c.items() if ".WIDE" in k)
        print(f, "total", sum(c.values()), "alu", alu, "fma", fma, "wide", wide, "| pipe cycles alu", 2 * alu, "fma", 2 * fma + 4 * wide)
        print("   ", dict(c.most_common(16)))
