#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one block per launch with the metrics the
DESIGN/bench rooflines quote plus the top warp-stall reasons.   usage: ncu_summary.py raw.csv"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',     # IMAD.WIDE issues here only: the binding pipe
        'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def main(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hdr, units = rows[0], rows[1]
    stall = [i for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    for r in rows[2:]:
        print('-----', r[hdr.index('Kernel Name')][:60])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]:>16s} {units[i]}")
        st = []
        for i in stall:
            try:
                st.append((float(r[i].replace(',', '')), hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
        st.sort(reverse=True)
        print('  stalls (warps per issue):', ', '.join(f"{n}={v:.2f}" for v, n in st[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
