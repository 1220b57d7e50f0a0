#!/usr/bin/env python3
"""Count SASS instructions per kernel (and per opcode) in a .so / .cubin: tools/sass_count.py <file> [filter]"""
import collections, re, subprocess, sys
out = subprocess.check_output(["cuobjdump", "-sass", sys.argv[1]], text=True)
flt = sys.argv[2] if len(sys.argv) > 2 else ""
f, cnt = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        f = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and f:
        cnt[f][m.group(2).split(".")[0] + (".WIDE" if ".WIDE" in m.group(2) else "")] += 1
for f, c in cnt.items():
    if flt in f:
        print(f, sum(c.values()), dict(c.most_common(14)))
