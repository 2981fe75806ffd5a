#!/usr/bin/env python
"""Per-kernel table from an `ncu --set full` report: duration, DRAM bytes and achieved GB/s, DRAM / L2 / SM throughput as % of
peak, achieved occupancy, registers.  One row per kernel name (first captured launch of each).
    python profiles/ncu_kernels.py gpurun_out/prof_stages_c3.ncu-rep"""
import csv
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "dram__bytes.sum.per_second": "bps",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed": "dram%",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit.lower(), 1)


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen = set()
    print("%-44s %9s %9s %9s %9s %6s %6s %6s %6s %5s" % ("kernel", "us", "rd MB", "wr MB", "GB/s", "dram%", "l2%", "sm%", "occ%", "regs"))
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        g = {}
        for m, k in WANT.items():
            if m in col:
                g[k] = (r[col[m]], units[col[m]])
        us = to_us(*g["dur"])
        if "rd" in g and g["rd"][0] not in ("", "n/a"):
            rd, wr = to_bytes(*g["rd"]), to_bytes(*g["wr"])
        else:   # reports captured without the full set carry the rate only: bytes = rate x duration
            v, unit = g.get("bps", ("0", "byte/s"))
            rate = float(v.replace(",", "")) * {"byte/s": 1, "kbyte/s": 1e3, "mbyte/s": 1e6, "gbyte/s": 1e9, "tbyte/s": 1e12}.get(unit.lower(), 1)
            rd, wr = rate * us * 1e-6, 0.0
        f = lambda k: float(g[k][0].replace(",", "")) if k in g and g[k][0] not in ("", "n/a") else float("nan")  # noqa: E731
        print("%-44s %9.1f %9.2f %9.2f %9.1f %6.1f %6.1f %6.1f %6.1f %5d" % (name[:44], us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3, f("dram%"), f("l2%"),
                                                                         f("sm%"), f("occ%"), int(f("regs"))))


if __name__ == "__main__":
    main()
