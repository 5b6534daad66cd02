#!/usr/bin/env python
"""Groups the per-instruction counts of tools/ncu_hot.py's listing into runs executed equally often.
usage: ncu_segments.py /tmp/hot_all.txt n_chunks"""
import re, sys
rows = []
for l in open(sys.argv[1]):
    m = re.match(r'\s*(\d+) (.{60}) inst\s+(\d+)', l)
    if m:
        rows.append((int(m.group(1)), m.group(2).strip(), int(m.group(3))))
N = float(sys.argv[2]) if len(sys.argv) > 2 else 524288.0
tot = 0
seg = None
def flush(seg):
    if seg:
        print("%5d-%5d  x%.2f/chunk  %4d instrs -> %.1f   %s" % (seg[0], seg[1], seg[2], seg[3], seg[2] * seg[3], seg[4]))
for n, s, c in rows:
    r = c / N
    if r < 0.02:
        continue
    tot += c
    if seg and abs(seg[2] - r) < 0.02 * max(1, r):
        seg[1] = n; seg[3] += 1
    else:
        flush(seg)
        seg = [n, n, r, 1, s[:50]]
flush(seg)
print("total per chunk", tot / N)
