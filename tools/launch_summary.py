#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [top]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        d[r[ki][:100]][0] += 1
        d[r[ki][:100]][1] += v
    tot = sum(v[1] for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%6d %10.1f us %5.1f%%  %s" % (v[0], v[1] / 1e3, 100 * v[1] / tot, k))
    print("total %.1f us over %d launches, %d distinct kernels" % (tot / 1e3, sum(v[0] for v in d.values()), len(d)))


if __name__ == "__main__":
    main()
