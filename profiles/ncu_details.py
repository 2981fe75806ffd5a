#!/usr/bin/env python
"""Summary of an ncu report: the Speed-of-Light / occupancy lines of the details page plus DRAM bytes, duration,
executed instructions and the per-issue stall reasons from the raw page.
    python profiles/ncu_details.py gpurun_out/prof.ncu-rep > profiles/<tag>_details.txt"""
import re
import subprocess
import sys

KEEP = ("Memory Throughput", "DRAM Throughput", "Duration", "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy", "L1/TEX Hit Rate",
        "L2 Hit Rate", "No Eligible", "Avg. Executed Instructions Per Scheduler", "Executed Instructions", "Registers Per Thread", "Waves Per SM",
        "Theoretical Occupancy", "Achieved Occupancy", "Avg. Active Threads Per Warp", "Shared Memory Configuration Size", "Static Shared Memory Per Block",
        "Block Limit Registers", "Block Limit Shared Mem")
RAW = ("dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__time_duration.sum", "smsp__inst_executed.sum ", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ")


def main():
    rep = sys.argv[1]
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    for l in det.splitlines():
        s = l.strip()
        if re.match(r"^\S.*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\)", s):
            print("  " + s)
        elif any(s.startswith(k) for k in KEEP):
            print("    " + re.sub(r"\s{2,}", "  ", s))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw"], capture_output=True, text=True).stdout
    stalls = []
    for l in raw.splitlines():
        s = re.sub(r"\s+", " ", l.strip())
        if any(s.startswith(k) for k in RAW):
            print(s)
        elif "average_warps_issue_stalled" in s and "per_issue_active" in s:
            p = s.split(" ")
            try:
                stalls.append((float(p[-1]), p[0]))
            except ValueError:
                pass
    for v, k in sorted(stalls, reverse=True)[:10]:
        print("%s %.3f" % (k, v))


if __name__ == "__main__":
    main()
