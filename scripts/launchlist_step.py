#!/usr/bin/env python
"""Developer tool: print the launches of the LAST generate() in an ncu launch list (scripts/gpu_launchlist.sh)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ki, vi, gi, bi, mi, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Grid Size', 'Block Size', 'Metric Name', 'ID'))
by = {}
for r in rows[start:]:
    by.setdefault(r[ii], {'name': r[ki], 'grid': r[gi], 'blk': r[bi]})[r[mi]] = r[vi]
L = [by[k] for k in sorted(by, key=int)]
idx = [i for i, r in enumerate(L) if 'rnn_tc_kernel' in r['name']]
# one generate() = everything after the previous step's post_proj, i.e. 2 launches after the previous LSTM... print between LSTMs
a, b = idx[-2], idx[-1]
tot = 0
fam = {}
for r in L[a + 1:b + 1]:
    n = r['name'].replace('ftb::', '').replace('void ', '')[:64]
    t = int(r['gpu__time_duration.sum'].replace(',', ''))
    print(f"{t / 1000:9.1f} us  {r['grid']:14s} {r['blk']:12s} tens={r.get('sm__inst_executed_pipe_tensor.sum', ''):>9s} {n}")
    tot += t
    k = n.split('(')[0]
    fam[k] = fam.get(k, 0) + t
print(f'{tot / 1e6:.3f} ms between two decoder-LSTM launches')
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
    print(f'{v / 1000:9.1f} us  {k}')
