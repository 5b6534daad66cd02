#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: instructions executed and stall samples
per SASS instruction, top entries and running totals.  usage: ncu_hot.py src.csv [kernel-index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split by kernels
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
hdr = blk[1]
ix = {h: i for i, h in enumerate(hdr)}
data = blk[2:]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(blk[0][1], "total warp-inst", tot_inst, "samples", tot_samp)
out = []
for n, r in enumerate(data):
    out.append((n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]),
                int(r[ix["stall_long_sb"]]), int(r[ix["stall_short_sb"]]), int(r[ix["stall_wait"]]),
                int(r[ix["stall_math"]]), int(r[ix["stall_not_selected"]])))
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
if mode == "all":
    for o in out:
        print("%5d %-60s inst %9d samp %6d long %5d short %5d wait %5d math %5d notsel %5d" % o)
else:
    for o in sorted(out, key=lambda o: -o[3])[:60]:
        print("%5d %-60s inst %9d samp %6d long %5d short %5d wait %5d math %5d notsel %5d" % o)
