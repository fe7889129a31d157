#!/usr/bin/env python
"""Top stall-sample SASS instructions of an ncu report (source page).  usage: ncu_hot.py rep [N] [--order PCT]
--order PCT: also list, in program order, every instruction holding at least PCT % of the samples."""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try: data.append((int(r[ci['# Samples']]), idx, r))
    except Exception: pass
tot = sum(d[0] for d in data) or 1
print(f'total samples {tot}, instructions {len(data)}')
agg = {}
for s, _, r in data:
    for h in stalls:
        try: agg[h] = agg.get(h, 0) + int(r[ci[h]])
        except Exception: pass
print('by reason: ' + ', '.join(f'{k[6:]} {100*v/tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, idx, r in sorted(data, key=lambda d: -d[0])[:n]:
    why = max(stalls, key=lambda h: int(r[ci[h]] or 0))
    print(f'{100*s/tot:5.1f}% #{idx:5d} {why[6:]:12s} {r[ci["Source"]].strip()[:80]}')
if '--order' in sys.argv:
    pct = float(sys.argv[sys.argv.index('--order') + 1])
    print(f'--- program order, >= {pct} % of samples')
    for s, idx, r in sorted(data, key=lambda d: d[1]):
        if 100 * s / tot >= pct:
            why = max(stalls, key=lambda h: int(r[ci[h]] or 0))
            print(f'{100*s/tot:5.1f}% #{idx:5d} {why[6:]:12s} {r[ci["Source"]].strip()[:90]}')
