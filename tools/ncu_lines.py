#!/usr/bin/env python
"""Developer tool: top CUDA source lines of one kernel by warp-stall samples and executed instructions, from
`ncu -i REP --page source --csv --print-source cuda,sass --kernel-name K > file.csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name",) or hdr is None or r[0] == "":
        continue
    try:
        smp = int(r[hdr.index("# Samples")] or 0)
        ins = int(r[hdr.index("Instructions Executed")] or 0)
        thr = int(r[hdr.index("Thread Instructions Executed")] or 0)
    except ValueError:
        continue
    out.append((smp, ins, thr, cur, r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print("samples %d  warp instructions %d  lane efficiency %.1f/32" % (tot, toti, sum(o[2] for o in out) / toti))
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% ins %4.1f thr  %s:%s  %s" % (100 * o[0] / tot, 100 * o[1] / toti, o[2] / max(o[1], 1), o[3], o[4], o[5]))
