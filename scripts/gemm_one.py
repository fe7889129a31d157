#!/usr/bin/env python
"""Developer tool for ncu: one GEMM shape, a few launches.  usage: gemm_one.py N Cin k out(16|32) [B S]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
from forwardtacotron_b200 import _lib
import importlib.util
spec = importlib.util.spec_from_file_location('gs', 'scripts/gemm_shapes.py')
N, Cin, k, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
B, S = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (64, 1245)
lib = _lib.lib()
dev = torch.device('cuda')
cp = (Cin + 63) // 64 * 64
w = torch.randn(N, Cin, k, device=dev) / (Cin * k) ** 0.5
wp = torch.empty(N * k * cp, dtype=torch.bfloat16, device=dev)
_lib.check(lib.ftb_pack_conv_weight(_lib.ptr(w), _lib.ptr(wp), N, Cin, k, N, cp, 1, None))
x = (torch.randn(B, S, cp, device=dev) * 0.5).bfloat16()
ldo = (N + 63) // 64 * 64
o = torch.empty(B, S, ldo, dtype=torch.bfloat16 if out == '16' else torch.float32, device=dev)
bias = torch.randn(N, device=dev)
d = _lib.ConvDesc()
d.B, d.S, d.Cin, d.N, d.ktaps, d.pad_left = B, S, cp, N, k, k // 2
d.lda, d.ldo, d.n_offset, d.relu = cp, ldo, 0, 0
d.bias = bias.data_ptr()
d.out_scale = 1.0
if out == '16':
    d.out_bf16 = o.data_ptr()
else:
    d.out_f32 = o.data_ptr()
for _ in range(8):
    _lib.check(lib.ftb_conv_gemm_bf16(_lib.ptr(x), _lib.ptr(wp), C.byref(d), None))
torch.cuda.synchronize()
print('done')
