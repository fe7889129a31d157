#!/usr/bin/env python
"""Developer tool: the tcgen05 GEMM on the layer shapes of cfg2 (frame rate: B 64 x S 1245; phoneme rate: S 200),
CUDA-event timings, best of 5 x 10 launches.  Run once with FTB_EPI_LEGACY=1 and once without to compare epilogues."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib

lib = _lib.lib()
dev = torch.device('cuda')
g = torch.Generator(device='cuda').manual_seed(0)


def pack(N, Cin, k):
    cp = (Cin + 63) // 64 * 64
    w = torch.randn(N, Cin, k, device=dev, generator=g) / (Cin * k) ** 0.5
    wp = torch.empty(N * k * cp, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.ftb_pack_conv_weight(_lib.ptr(w), _lib.ptr(wp), N, Cin, k, N, cp, 1, None))
    return wp, cp


def desc(B, S, cp, N, k, ldo, n_off, relu, bias, bn, out16=None, out32=None, out_t=None, res16=None, ldr=0):
    d = _lib.ConvDesc()
    d.B, d.S, d.Cin, d.N, d.ktaps, d.pad_left = B, S, cp, N, k, k // 2
    d.lda, d.ldo, d.n_offset, d.relu = cp, ldo, n_off, int(relu)
    keep = []
    if bias:
        t = torch.randn(N, device=dev); keep.append(t); d.bias = t.data_ptr()
    if bn:
        a, b = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev); keep += [a, b]; d.scale, d.shift = a.data_ptr(), b.data_ptr()
    d.out_scale = 1.0
    if out16 is not None: d.out_bf16 = out16.data_ptr()
    if out32 is not None: d.out_f32 = out32.data_ptr()
    if out_t is not None: d.out_t = out_t.data_ptr()
    if res16 is not None: d.residual_bf16, d.ldr = res16.data_ptr(), ldr
    return d, keep


def timeit(fn):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    return best * 1e3


def single(name, B, S, Cin, N, k, relu, bias, bn, out, res=False):
    wp, cp = pack(N, Cin, k)
    x = (torch.randn(B, S, cp, device=dev, generator=g) * 0.5).bfloat16()
    ldo = (N + 63) // 64 * 64
    o16 = torch.empty(B, S, ldo, dtype=torch.bfloat16, device=dev) if '16' in out else None
    o32 = torch.empty(B, S, ldo, dtype=torch.float32, device=dev) if '32' in out else None
    ot = torch.empty(B, N, S, dtype=torch.float32, device=dev) if 't' in out else None
    r16 = torch.randn(B, S, ldo, device=dev).bfloat16() if res else None
    d, keep = desc(B, S, cp, N, k, ldo, 0, relu, bias, bn, o16, o32, ot, r16, ldo)
    us = timeit(lambda: _lib.check(lib.ftb_conv_gemm_bf16(_lib.ptr(x), _lib.ptr(wp), C.byref(d), None)))
    fl = 2.0 * B * S * N * k * Cin
    print(f'{name:34s} {us:8.1f} us  {fl / us / 1e6:7.0f} TFLOP/s')


def bank(name, B, S, Cin, ch, K):
    cp = (Cin + 63) // 64 * 64
    x = (torch.randn(B, S, cp, device=dev, generator=g) * 0.5).bfloat16()
    out = torch.empty(B, S, K * ch, dtype=torch.bfloat16, device=dev)
    descs = (_lib.ConvDesc * K)(); wptrs = (C.c_void_p * K)(); keep = []
    fl = 0.0
    for i in range(K):
        k = i + 1
        wp, _ = pack(ch, Cin, k)
        d, kp = desc(B, S, cp, ch, k, K * ch, i * ch, True, False, True, out)
        descs[i] = d; wptrs[i] = wp.data_ptr(); keep += [wp, kp]
        fl += 2.0 * B * S * ch * k * Cin
    us = timeit(lambda: _lib.check(lib.ftb_conv_bank_bf16(_lib.ptr(x), wptrs, descs, K, 1, None)))
    print(f'{name:34s} {us:8.1f} us  {fl / us / 1e6:7.0f} TFLOP/s')


print('FTB_EPI_LEGACY =', os.environ.get('FTB_EPI_LEGACY', '0'))
single('k1 256->512 bias, bf16 out (L)', 64, 1245, 256, 512, 1, False, True, False, '16')
single('k1 256->1536 bias, f32 out (L)', 64, 1245, 256, 1536, 1, False, True, False, '32')
single('k1 512->4096 bias, f32 out (T)', 64, 200, 512, 4096, 1, False, True, False, '32')
single('k3 2048->256 relu bn, bf16 (L)', 64, 1245, 2048, 256, 3, True, False, True, '16')
single('k3 256->80 bn + res, bf16 (L)', 64, 1245, 256, 80, 3, False, False, True, '16', res=True)
single('k1 128->256, bf16 (L)', 64, 1245, 128, 256, 1, False, False, False, '16')
single('k1 1024->80 bias, out_t+bf16 (L)', 64, 1245, 1024, 80, 1, False, True, False, 't16')
single('k3 4096->256 relu bn, bf16 (T)', 64, 200, 4096, 256, 3, True, False, True, '16')
bank('bank 80(128) x8 -> 2048 pool (L)', 64, 1245, 80, 256, 8)
bank('bank 256 x16 -> 4096 pool (T)', 64, 200, 256, 256, 16)
