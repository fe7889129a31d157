#!/usr/bin/env python
"""Developer tool: one-line digest of a bench.py JSON line (ms/step, per-family kernel times, extras)."""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"ms/step {d['ms_per_step']:.3f}  e2e {d.get('e2e', {}).get('ms_per_step', 0):.3f}  lat1 {d.get('latency_ms_in_flight_1', 0):.2f}  "
      + '  '.join(f"{k['name']} {k['ms_per_step']:.3f}" for k in d.get('kernels', [])))
ex = d.get('extra', {})
if 'fast_pitch' in ex:
    print('  fp', round(ex['fast_pitch']['ms_per_step'], 2), [(k['name'], round(k['ms_per_step'], 2)) for k in ex['fast_pitch'].get('kernels', [])])
if 'stft_mel' in ex:
    print('  stft', round(ex['stft_mel']['value'] / 1e6, 2), 'M audio-s/s', ex['stft_mel'].get('roofline', {}).get('frac'))
if 'long_article' in ex:
    print('  long', {k: round(v['frames_per_s'] / 1e6, 2) for k, v in ex['long_article'].items() if isinstance(v, dict)})
