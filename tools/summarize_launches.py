#!/usr/bin/env python3
"""Summarise an `ncu --csv --metrics ...` launch list per kernel (mean per launch)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        d = dict(zip(h, r))
        k = d["Kernel Name"].split("(")[0]
        agg.setdefault(k, collections.OrderedDict()).setdefault(d["Metric Name"], []).append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(m.get("gpu__time_duration.sum", [0])) for m in agg.values())
    for k, m in agg.items():
        t = m.get("gpu__time_duration.sum", [0])
        print("%-14s n=%2d  %8.1f us  share %5.1f%%" % (k, len(t), sum(t) / len(t) / 1000.0, 100.0 * sum(t) / tot), end="")
        for n, v in m.items():
            if n != "gpu__time_duration.sum":
                short = {"smsp__thread_inst_executed_per_inst_executed.ratio": "lanes", "smsp__inst_executed.sum": "winst(M)",
                         "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%", "dram__bytes_read.sum": "rdMB",
                         "dram__bytes_write.sum": "wrMB"}.get(n, n)
                x = sum(v) / len(v)
                if short in ("winst(M)", "rdMB", "wrMB"):
                    x /= 1e6
                print("  %s %.1f" % (short, x), end="")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
