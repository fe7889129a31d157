#!/usr/bin/env python
"""Developer tool: per-step phase clocks of the ping-pong decoder-LSTM kernel (CTA 0 of cluster 0)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
fn = C.CDLL(str(_lib.lib_path())).ftb_debug_rnn_timing
fn.argtypes = [C.c_void_p]
H, B, S = 512, 64, 300
xg = torch.randn(B, S, 2, 4 * H, device='cuda') * 0.3
whh = torch.randn(2, 4 * H, H, device='cuda') / H ** 0.5
out = torch.empty(B, S, 2 * H, device='cuda')
dbg = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
for it in range(2):
    fn(dbg.data_ptr() if it else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.ftb_rnn_bidir(_lib.ptr(xg), _lib.ptr(whh), None, _lib.ptr(out), B, S, H, 1, 0, _lib.current_stream(out.device)))
    e1.record()
    torch.cuda.synchronize()
    print(f'B={B} S={S}: {e0.elapsed_time(e1) * 1e3 / S:.2f} us/step')
fn(None)
d = dbg.cpu().view(64, 16)[20:60].double()
step = (d[1:, 0] - d[:-1, 0]).mean()
print(f'clocks per step {step:.0f}')
names = {15: 'gate loop top (c=1 entry)', 0: 'A d_full', 1: 'A tmem read', 2: 'A gates+h', 3: 'A bar.sync', 4: 'A pushed',
         5: 'B d_full', 6: 'B tmem read', 7: 'B gates+h', 8: 'B bar.sync', 9: 'B pushed',
         10: 'iss A own_ready', 11: 'iss A all issued', 12: 'iss A commit', 13: 'iss B own_ready', 14: 'iss B all issued'}
base = d[:, 0]
for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14):
    print(f'  {names[k]:28s} @ {(d[:, k] - base).mean():8.0f}')
print(f'  next A d_full              @ {step:8.0f}')
