#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total us, share.
    python profiles/launch_summary.py gpurun_out/launches.csv [skip_first_n_launches]"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lines = [l for l in open(path, errors="ignore") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    iN, iV, iU, iID = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    agg = OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= iV or not r[iID].isdigit() or int(r[iID]) < skip:
            continue
        v = float(r[iV].replace(",", ""))
        v = v / 1000.0 if r[iU].startswith("ns") else (v * 1000.0 if r[iU].startswith("ms") else v)
        name = r[iN].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("%d launches, %.1f us total (per-launch times are cold-cache and serialised: compare shares)" % (n, tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%7.2f%% %10.1f us %5d x  %s" % (100 * a[1] / tot, a[1], a[0], k))


if __name__ == "__main__":
    main()
