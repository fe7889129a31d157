#!/usr/bin/env python
"""Developer tool (needs a build with FTB_NVCC_DEFINES=FTB_PHASE_TIMING: rm -rf forwardtacotron_b200/csrc/build && FTB_NVCC_DEFINES=FTB_PHASE_TIMING python __graft_entry__.py build): per-step phase clocks of the tcgen05 recurrence kernel (CTA 0 of cluster 0)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
fn = C.CDLL(str(_lib.lib_path())).ftb_debug_rnn_timing
fn.argtypes = [C.c_void_p]
NAMES = ['h landed', 'mma issued', 'acc ready', 'tmem read', 'gates done', 'fence+bar', 'pushed']
for (H, lstm, B, S) in [(512, 1, 64, 200), (512, 1, 16, 200)]:
    G = 4 if lstm else 3
    xg = torch.randn(B, S, 2, G * H, device='cuda') * 0.3
    whh = torch.randn(2, G * H, H, device='cuda') / H ** 0.5
    bhn = torch.zeros(2, H, device='cuda')
    out = torch.empty(B, S, 2 * H, device='cuda')
    dbg = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
    for it in range(2):
        fn(dbg.data_ptr() if it else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.ftb_rnn_bidir(_lib.ptr(xg), _lib.ptr(whh), None if lstm else _lib.ptr(bhn), _lib.ptr(out), B, S, H,
                                     lstm, 0, _lib.current_stream(out.device)))
        e1.record()
        torch.cuda.synchronize()
        print(f'H={H} lstm={lstm} B={B} S={S}: {e0.elapsed_time(e1) * 1e3 / S:.2f} us/step')
    fn(None)
    d = dbg.cpu().view(64, 8)
    st = d[20:60]
    print(f'  clocks/step {(st[1:, 2] - st[:-1, 2]).float().mean():.0f}')
    names = {3: 'tmem read', 4: 'gates done', 5: 'fence+bar', 6: 'pushed', 7: 'waited+mma issued'}
    for k in range(3, 8):
        print(f'  {names[k]:18s} +{(st[:, k] - st[:, k - 1]).float().mean():8.0f}')
    print(f'  {"acc ready (next)":18s} +{(st[1:, 2] - st[:-1, 7]).float().mean():8.0f}')
