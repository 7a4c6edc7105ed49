#!/usr/bin/env python
"""Hot SASS of the first kernel in an .ncu-rep: per-instruction executed counts and stall samples (read here, no GPU).
   python tools/ncu_hot.py rep [min_share_percent]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blk = rows[start[0] + 1:(start[1] if len(start) > 1 else len(rows))]
H = blk[0]
ia, isrc, ie, iss = H.index("Address"), H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
tot = sum(int(r[ie] or 0) for r in blk[1:])
tots = sum(int(r[iss] or 0) for r in blk[1:])
print("total warp instructions", tot, "samples", tots)
for n, r in enumerate(blk[1:]):
    e, s = int(r[ie] or 0), int(r[iss] or 0)
    if 100.0 * e / tot >= thr or 100.0 * s / max(tots, 1) >= thr:
        print("%5d %6.2f%% exec %6.2f%% stall  %s" % (n, 100.0 * e / tot, 100.0 * s / max(tots, 1), r[isrc].strip()))
