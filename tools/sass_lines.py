#!/usr/bin/env python
"""Attribute SASS instructions to source lines (development aid).
    cuobjdump -xelf all x.o; nvdisasm -g -c x.cubin > x.dis; python tools/sass_lines.py x.dis [file-substring] [lo-hi]
Prints instructions per source line (static count), optionally restricted to one source file and a line range."""
import collections
import re
import sys

def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    lo, hi = (0, 10**9)
    if len(sys.argv) > 3:
        lo, hi = [int(x) for x in sys.argv[3].split("-")]
    cur = None
    cnt = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for ln in open(path):
        m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(2).split(".")[0]] += 1
    tot = 0
    for (f, l), n in sorted(cnt.items()):
        if sub in f and lo <= l <= hi:
            tot += n
            print(f"{f}:{l:4d} {n:4d}  " + " ".join(f"{k}{v}" for k, v in ops[(f, l)].most_common(6)))
    print("total", tot)

if __name__ == "__main__":
    main()
