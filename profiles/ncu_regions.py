#!/usr/bin/env python
"""Group the per-line instruction shares of an ncu report by named source regions of raster.cu's fine_warp_k.
    python profiles/ncu_regions.py gpurun_out/prof.ncu-rep   (regions are found from marker comments in the current source)"""
import csv, io, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "vkvg_b200", "csrc", "raster.cu")).read().split("\n")
def line_of(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + pat)
marks = [
    ("row_edge_masks (rows = lanes)", line_of("void row_edge_masks(")),
    ("plane_add", line_of("void plane_add(")),
    ("blend_int", line_of("uint32_t blend_int(")),
    ("(fine_build_group/fine_k...)", line_of("// Warp 0 turns the next path-tiles")),
    ("fw_chunk: V term", line_of("void fw_chunk(")),
    ("fw_chunk: row ranges + scan", line_of("// ---- H term: one work item per (edge, spanned sample row) ----")),
    ("fw_chunk: item search + shuffles", line_of("for (int t0 = 0; t0 < total; t0 += 32) {")),
    ("fw_chunk: item crossing + atomic", line_of("if (t < total) {")),
    ("(fw_chunk_far)", line_of("void fw_chunk_far(")),
    ("kernel prologue / tile load", line_of("fine_warp_k(FineArgs a, uint32_t n_tiles) {")),
    ("path-tile header + paint setup", line_of("for (uint32_t pb = first; pb < end; pb += 32)")),
    ("short lists: plane loop", line_of("// ---- short lists: winding in bit-sliced counters")),
    ("long lists: item loop + extraction", line_of("// ---- long lists: (edge, sample row) work items")),
    ("queue build", line_of("if (skip) continue;", line_of("// ---- long lists: (edge, sample row) work items"))),
    ("blend rounds (bookkeeping)", line_of("// ---- blend, 32 pixels at a time ----")),
    ("resolve", line_of("// ---- resolve; pixels whose samples differ")),
    ("(after kernel)", line_of("// Which kernel serves a batch without clip state")),
]
BO0, BO1 = line_of("uint32_t unorm8(float v)"), line_of("uint32_t blend_over(") + 10
EG0, EG1 = line_of("void eval_gradient("), line_of("uint32_t unorm8(float v)") - 2
def region(n):
    name = "(before)"
    for nm, l in marks:
        if n >= l:
            name = nm
    return name
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
per, fname, hdr = {}, None, None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r; iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit() and r[2] == "-":
        n, s = int(r[iI] or 0), int(r[iS] or 0)
        key = region(int(r[0])) if fname == "raster.cu" else "other: " + fname
        if fname == "raster.cu" and BO0 <= int(r[0]) <= BO1: key = "blend_over / unorm8 (fp32)"
        if fname == "raster.cu" and EG0 <= int(r[0]) < EG1: key = "paint evaluation (gradient / surface)"
        a = per.setdefault(key, [0, 0]); a[0] += n; a[1] += s
ti = sum(a[0] for a in per.values()) or 1; ts = sum(a[1] for a in per.values()) or 1
print("total warp instructions %d" % ti)
for k, a in sorted(per.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f%% inst %6.2f%% samples  %s" % (100.0 * a[0] / ti, 100.0 * a[1] / ts, k))
