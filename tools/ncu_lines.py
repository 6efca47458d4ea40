#!/usr/bin/env python3
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump per CUDA source line:
warp-level instructions executed, stall samples, average active threads.  usage: ncu_lines.py dump.csv [top]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname = "?"
    agg = {}
    h = None
    cur = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            h = r
            ie = h.index("Instructions Executed"); ns = h.index("# Samples"); te = h.index("Thread Instructions Executed")
            continue
        if h is None or len(r) < 8:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip()[:100])
            agg.setdefault(cur, [0, 0, 0])
            continue
        if cur is None:
            continue
        try:
            a = agg[cur]
            a[0] += int(r[ie]); a[1] += int(r[ns]); a[2] += int(r[te])
        except ValueError:
            pass
    tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
    print("total warp-instr %d, samples %d" % (tot, tots))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.1f%% inst %5.1f%% stall  lanes %4.1f  %s:%d  %s" % (100.0 * a[0] / max(tot, 1), 100.0 * a[1] / max(tots, 1), a[2] / max(a[0], 1), k[0], k[1], k[2]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
