#!/usr/bin/env python
"""Summarise an `ncu --set full` report: python tools/ncu_summary.py report.ncu-rep > profiles/xxx_summary.txt
(reads `ncu -i ... --page raw --csv`; one line per captured launch with the counters DESIGN.md argues from)."""
import csv
import io
import json
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), short) for k, short in WANT if k in hdr]
    print("# " + " | ".join(f"{s}[{units[i]}]" for i, s in cols))
    js = {}
    for r in rows[2:]:
        vals = []
        for i, s in cols:
            v = r[i]
            if s == "kernel":
                v = v.replace("void <unnamed>::", "").split("(")[0][:34]
            else:
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
            vals.append(v)
        print(" | ".join(vals))
        name = vals[0]
        rd, wr = r[hdr.index("dram__bytes_read.sum")], r[hdr.index("dram__bytes_write.sum")]
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        js.setdefault(name, {"dram_bytes": float(rd) * mult.get(ur, 1) + float(wr) * mult.get(uw, 1),
                             "ms": float(r[hdr.index("gpu__time_duration.sum")])})
    if len(sys.argv) > 2:
        json.dump(js, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
