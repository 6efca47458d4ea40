#!/usr/bin/env python3
"""Per-kernel SASS profile from `ncu -i X.ncu-rep --page source --csv`: instruction share / lanes per address segment."""
import csv
import sys


def main(path, nseg=24):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1][:60], "rows": []}
            blocks.append(cur)
            continue
        if cur is not None:
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if b["name"] in seen:
            continue
        seen.add(b["name"])
        hdr = b["rows"][0]
        ci = {n: i for i, n in enumerate(hdr)}
        data = [(r[ci["Source"]].strip(), int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]), float(r[ci["Avg. Threads Executed"]] or 0))
                for r in b["rows"][1:] if len(r) > 10]
        tot = sum(d[1] for d in data) or 1
        toti = sum(d[2] for d in data) or 1
        print("==", b["name"], "sass", len(data), "winst", toti)
        n = len(data)
        for k in range(nseg):
            seg = data[k * n // nseg:(k + 1) * n // nseg]
            s = sum(d[1] for d in seg); i = sum(d[2] for d in seg)
            thr = sum(d[3] * d[2] for d in seg) / max(1, i)
            print("seg %2d idx %4d-%4d samples %5.1f%% inst %5.1f%% lanes %4.1f  %s" % (k, k * n // nseg, (k + 1) * n // nseg, 100 * s / tot, 100 * i / toti, thr, seg[0][0][:48]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
