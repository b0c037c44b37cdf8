#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path, errors="replace")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
            k = d["Kernel Name"].split("(")[0]
            agg[k][0] += 1
            agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:40s} {v[0]:6d} launches {v[1]:10.3f} ms {100 * v[1] / tot:5.1f}%")
    print(f"{'total':40s} {sum(v[0] for v in agg.values()):6d} launches {tot:10.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1])
