#!/usr/bin/env python
"""Developer tool: the tcgen05 attention alone on the cfg3 postnet shape (B 128 x S 1954, 2 heads x 128), CUDA events,
best of 5 x 10 launches; also the pitch / energy predictor shape (hd 64)."""
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
dev = torch.device('cuda')
for (B, S, E, heads, masked) in [(128, 1954, 256, 2, True), (128, 1954, 256, 2, False), (128, 300, 128, 2, True)]:
    qkv = (torch.randn(B, S, 3 * E, device=dev) * 0.5).half()
    tok = torch.ones(B, S, dtype=torch.int64, device=dev)
    if masked:
        for b in range(B):
            tok[b, S - (b * 7) % (S // 8):] = 0
    ctx = torch.empty(B, S, E, dtype=torch.float16, device=dev)
    fn = lambda: _lib.check(lib.ftb_attention_16(_lib.ptr(qkv), _lib.ptr(tok), _lib.ptr(ctx), B, S, E, heads, 1, 0,
                                                  _lib.current_stream(dev)))
    for _ in range(3): fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    fl = 4.0 * B * S * S * E
    print(f'B{B} S{S} E{E} h{heads} masked={masked}: {best * 1e3:8.1f} us  {fl / best / 1e9:6.0f} TFLOP/s')
