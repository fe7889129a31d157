#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep, --set full) into one block per kernel launch: the numbers DESIGN.md and
bench.py's roofline cite.   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x.txt"""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid (CTAs)'),
    ('launch__block_size', 'block'),
    ('launch__cluster_size', 'cluster size'),
    ('launch__cluster_max_active', 'max active clusters'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__waves_per_multiprocessor', 'waves/SM'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__bytes_read.sum.per_second', 'DRAM read rate'),
    ('dram__bytes_write.sum.per_second', 'DRAM write rate'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe % (elapsed)'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__cycles_elapsed.max', 'cycles'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
]
STALLS = 'smsp__average_warps_issue_stalled_'


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f'# ncu --set full summary of {path} (cold-cache, serialised replays: compare shares, not absolutes)')
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']][:110]}")
        for k, label in KEYS:
            if k in col and r[col[k]] not in ('', 'n/a'):
                print(f'  {label:28s} {r[col[k]]} {units[col[k]]}')
        stalls = []
        for h, i in col.items():
            if h.startswith(STALLS) and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try:
                    stalls.append((float(r[i]), h[len(STALLS):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        top = ', '.join(f'{n} {v:.2f}' for v, n in sorted(stalls, reverse=True)[:6])
        print(f'  {"top stalls (warps/issue)":28s} {top}')


if __name__ == '__main__':
    main(sys.argv[1])
