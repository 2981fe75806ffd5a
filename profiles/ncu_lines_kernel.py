import csv, io, subprocess, sys
rep = sys.argv[1]; want = int(sys.argv[2]); top = int(sys.argv[3])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; fname = None; hdr = None; per = {}; idx = 0
for r in rows:
    if not r: continue
    if r[0] == "Function Name": idx += 1
    elif r[0] == "File Path": fname = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r; iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit() and r[2] == "-" and idx == want:
        n, s = int(r[iI] or 0), int(r[iS] or 0)
        if n or s: per[(fname, int(r[0]), r[1].strip()[:120])] = [n, s]
tot = sum(a[0] for a in per.values()) or 1; ts = sum(a[1] for a in per.values()) or 1
print("warp instructions", tot, "samples", ts)
for k, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% %6.2f%%  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / ts, k[0], k[1], k[2]))
