#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of executed warp instructions and of stall samples.
    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
(runs `ncu -i <rep> --page source --print-source cuda,sass --csv`; kernels must be compiled with -lineinfo)"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    per = {}
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            print("kernel:", r[1])
        elif r[0] == "Line No":
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and r[0].isdigit() and r[2] == "-":   # the per-line aggregate row (SASS rows carry an address)
            n, s = int(r[iI] or 0), int(r[iS] or 0)
            if n or s:
                k = (fname, int(r[0]))
                a = per.setdefault(k, [0, 0, r[1]])
                a[0] += n
                a[1] += s
    tot_i = sum(a[0] for a in per.values()) or 1
    tot_s = sum(a[1] for a in per.values()) or 1
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%7s %7s  %s" % ("inst%", "smpl%", "line"))
    for (f, l), a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%% %6.2f%%  %s:%d  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, f, l, a[2].strip()[:100]))


if __name__ == "__main__":
    main()
