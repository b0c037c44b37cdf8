#!/usr/bin/env python3
"""profiles/r02_traffic.json from an `ncu --set full` capture of bench.py's dominant kernel.

    ncu --set full --clock-control none -k regex:k_msm_accumulate --launch-skip <launches of setup + warm-up> -c 7 -o gpurun_out/bench_acc \\
        python bench.py --streams 1 --steps 2 --warmup 3 --no-cpu-baseline
    ncu -i gpurun_out/bench_acc.ncu-rep --page raw --csv > gpurun_out/bench_acc_raw.csv
    python tools/ncu_traffic.py gpurun_out/bench_acc_raw.csv

The 7 captured launches are the 7 commitment rounds of ONE proof, so the average is per launch over a
proof exactly like bench.py's `algorithmic_bytes_per_launch`."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    return float(x.replace(",", ""))


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    ih = hdr.index("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = n = 0
    heavy = dur = 0.0
    for r in rows[2:]:
        if "k_msm_accumulate" not in r[hdr.index("Kernel Name")]:
            continue
        tot += num(r[ir]) * scale[units[ir]] + num(r[iw]) * scale[units[iw]]
        heavy += num(r[ih]) * num(r[it])
        dur += num(r[it])
        n += 1
    out = {"k_msm_accumulate_dram_bytes_per_launch": tot / n, "launches": n,
           "k_msm_accumulate_fmaheavy_pct_time_weighted": heavy / dur,
           "source": "ncu --set full --clock-control none, k_msm_accumulate launches of one proof of "
                     "`bench.py --streams 1` (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch); "
                     "raw export under profiles/"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main(sys.argv[1])
