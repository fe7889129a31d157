#!/usr/bin/env python
"""Developer tool: decoder-LSTM recurrence at the cfg2 shape (B 64, S 1245) with / without the frame -> phoneme row
index and the hi/lo output pair, at the two cluster chunk sizes.  us/step from CUDA events, best of 3."""
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
B, T, S, H, G = 64, 200, 1245, 512, 4
g = torch.Generator().manual_seed(0)
rows = (torch.randn(B * T + 1, 2, G * H, generator=g) * 0.3).cuda()
reps = torch.randint(4, 9, (B, T), generator=g)
idx = torch.full((B, S), B * T, dtype=torch.int32)
for b in range(B):
    seq = torch.repeat_interleave(torch.arange(T) + b * T, reps[b])[:S]
    idx[b, :len(seq)] = seq.int()
idx = idx.cuda()
dense = rows[idx.long()].contiguous()
whh = (torch.randn(2, G * H, H, generator=g) / H ** 0.5).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def run(xg, xrow, kind, ldo, lo_off):
    out = torch.empty(B, S, max(ldo, 2 * H), dtype=(torch.float32, torch.bfloat16, torch.float16)[kind], device='cuda')
    best = 1e9
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ftb_rnn_bidir_rows(_lib.ptr(xg), _lib.ptr(xrow) if xrow is not None else None, _lib.ptr(whh), None,
                                          _lib.ptr(out), B, S, H, 1, kind, ldo, lo_off, _lib.current_stream(out.device)))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for chunk in (24, 32):
    _lib.check(lib.ftb_tune(_lib.FTB_TUNE_LSTM_MIN_CHUNK, chunk))
    for name, xg, xrow, kind, ldo, lo in (('dense f16 plain', dense, None, 2, 0, 0), ('dense f16 hi/lo', dense, None, 2, 4 * H, 2 * H),
                                          ('rows  f16 plain', rows, idx, 2, 0, 0), ('rows  f16 hi/lo', rows, idx, 2, 4 * H, 2 * H),
                                          ('dense bf16 plain', dense, None, 1, 0, 0), ('rows  f32 out', rows, idx, 0, 0, 0)):
        ms = run(xg, xrow, kind, ldo, lo)
        print(f'chunk {chunk} {name:18s} {ms:7.3f} ms  {ms * 1e3 / S:.3f} us/step')
