#!/usr/bin/env python
"""Developer tool (needs a build with FTB_NVCC_DEFINES=FTB_PHASE_TIMING: rm -rf forwardtacotron_b200/csrc/build && FTB_NVCC_DEFINES=FTB_PHASE_TIMING python __graft_entry__.py build): per-tile phase clocks of one softmax thread of the tcgen05 attention kernel (CTA 0)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
fn = C.CDLL(str(_lib.lib_path())).ftb_debug_attn_timing
fn.argtypes = [C.c_void_p]
dev = torch.device('cuda')
B, S, E, heads = 128, 1954, 256, 2
qkv = (torch.randn(B, S, 3 * E, device=dev) * 0.5).half()
ctx = torch.empty(B, S, E, dtype=torch.float16, device=dev)
dbg = torch.zeros(16 * 10 + 16, dtype=torch.int64, device=dev)
for it in range(3):
    fn(dbg.data_ptr() if it == 2 else None)
    _lib.check(lib.ftb_attention_16(_lib.ptr(qkv), None, _lib.ptr(ctx), B, S, E, heads, 1, 0, _lib.current_stream(dev)))
    torch.cuda.synchronize()
fn(None)
x = dbg.cpu()
d = x[:160].view(16, 10)
print('CTA 0: entry -> softmax loop', int(x[161] - x[160]), '| loop (16 tiles)', int(x[162] - x[161]), '| normalise + store', int(x[163] - x[162]), '| final sync', int(x[164] - x[163]), '| first tile', int(d[1, 0] - d[0, 0]))
names = ['mask+bar', 'wait S', 'pass 1', 'pair bar', 'rescale chk', 'pass-2 loads', 'exp + pack', 'P stores', 'arrive']
print('clocks per tile', float((d[5:15, 0] - d[4:14, 0]).float().mean()))
for k in range(1, 10):
    print(f'  {names[k - 1]:14s} +{float((d[4:15, k] - d[4:15, k - 1]).float().mean()):8.0f}')
print(f'  {"loop back":14s} +{float((d[5:15, 0] - d[4:14, 9]).float().mean()):8.0f}')
