#!/usr/bin/env python
"""Developer tool (needs a build with FTB_NVCC_DEFINES=FTB_PHASE_TIMING: rm -rf forwardtacotron_b200/csrc/build && FTB_NVCC_DEFINES=FTB_PHASE_TIMING python __graft_entry__.py build): per-step phase clocks of the GRU-256 cluster kernel (CTA 0 of cluster 0)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
fn = C.CDLL(str(_lib.lib_path())).ftb_debug_gru_timing
fn.argtypes = [C.c_void_p]
H, B, S = 256, 64, 400
xg = torch.randn(B, S, 2, 3 * H, device='cuda') * 0.3
whh = torch.randn(2, 3 * H, H, device='cuda') / H ** 0.5
bhn = torch.zeros(2, H, device='cuda')
out = torch.empty(B, S, 2 * H, dtype=torch.float16, device='cuda')
dbg = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
for it in range(2):
    fn(dbg.data_ptr() if it else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.ftb_rnn_bidir(_lib.ptr(xg), _lib.ptr(whh), _lib.ptr(bhn), _lib.ptr(out), B, S, H, 0, 2, _lib.current_stream(out.device)))
    e1.record()
    torch.cuda.synchronize()
    print(f'GRU H={H} B={B} S={S}: {e0.elapsed_time(e1) * 1e3 / S:.3f} us/step')
fn(None)
st = dbg.cpu().view(64, 8)[20:60]
print(f'  clocks/step {(st[1:, 0] - st[:-1, 0]).float().mean():.0f}')
names = ['h landed', 'MMAs', 'acc->smem + barrier', 'gate maths', 'barrier', 'push']
for k in range(1, 7):
    print(f'  {names[k - 1]:22s} +{(st[:, k] - st[:, k - 1]).float().mean():7.0f}')
print(f'  {"stores + loop":22s} +{(st[1:, 0] - st[:-1, 6]).float().mean():7.0f}')
