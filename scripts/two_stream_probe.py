#!/usr/bin/env python
"""Probe: throughput of K generate() calls issued round-robin on S CUDA streams (S model replicas) vs one stream."""
import sys, time
sys.path.insert(0, '.')
import torch
from forwardtacotron_b200.utils import synth

dev = torch.device('cuda', 0)
K = 12
for S in (1, 2, 3):
    models, xs, streams = [], [], []
    for i in range(S):
        m, _ = synth.synthetic_model('forward_tacotron')
        models.append(m.to(dev))
        xs.append(synth.synthetic_tokens(64, 200, seed=1 + i).to(dev))
        streams.append(torch.cuda.Stream(dev))
    for i in range(S):          # warm-up
        with torch.cuda.stream(streams[i]):
            for _ in range(3):
                out = models[i].generate(xs[i])
    torch.cuda.synchronize()
    frames = int(out['mel_len'].sum())
    t0 = time.perf_counter()
    for k in range(K):
        with torch.cuda.stream(streams[k % S]):
            out = models[k % S].generate(xs[k % S])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'streams {S}: {dt / K * 1e3:.2f} ms per generate, {frames * K / dt / 1e6:.2f} M frames/s')
