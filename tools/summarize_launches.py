#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: summarize_launches.py launches.csv [first_kernel_regex_of_a_step]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        rows.append((x["Kernel Name"], v, x["Grid Size"], x["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    marker = sys.argv[2] if len(sys.argv) > 2 else None
    if marker:
        idx = [i for i, r in enumerate(rows) if re.search(marker, r[0])]
        if len(idx) >= 2:
            rows = rows[idx[-2]:idx[-1]]
            print(f"# one step = launches between the last two '{marker}' ({len(rows)} launches)")
    tot = collections.OrderedDict()
    for name, v, g, b in rows:
        key = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")[:70]
        t = tot.setdefault(key, [0, 0.0, g, b])
        t[0] += 1
        t[1] += v
    s = sum(t[1] for t in tot.values())
    print(f"{'kernel':72s} {'n':>4s} {'us':>9s} {'us/launch':>9s} {'share':>6s}  grid block")
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {t[0]:4d} {t[1]:9.1f} {t[1]/t[0]:9.2f} {100*t[1]/s:5.1f}%  {t[2]} {t[3]}")
    print(f"{'total':72s} {sum(t[0] for t in tot.values()):4d} {s:9.1f}")


if __name__ == "__main__":
    main()
