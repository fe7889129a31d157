"""Implicit-GEMM conv1d kernels (fp32 SIMT and bf16 tcgen05) against torch's conv1d on the CPU."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from forwardtacotron_b200 import _lib

pytestmark = pytest.mark.gpu


def reference(x, w, k, relu, scale, shift, bias, res, out_scale):
    """x (B,S,Cin) -> (B,S,N): conv (pad k//2, truncated to S) + the epilogue order of ftb_conv_desc."""
    S = x.shape[1]
    y = F.conv1d(x.transpose(1, 2), w, None, 1, k // 2)[:, :, :S].transpose(1, 2)
    if bias is not None:
        y = y + bias
    if relu:
        y = torch.relu(y)
    if scale is not None:
        y = y * scale + shift
    if res is not None:
        y = y + res
    return y * out_scale


def run_kernel(x, w, k, relu, scale, shift, bias, res, out_scale, bf16, want_t=False, n_offset=0, extra_cols=0):
    lib = _lib.lib()
    B, S, Cin = x.shape
    N = w.shape[0]
    dev = torch.device('cuda')
    cin_pad = (Cin + 63) // 64 * 64
    wdev = w.to(dev).contiguous()
    wp = torch.empty(N * k * cin_pad, dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
    _lib.check(lib.ftb_pack_conv_weight(_lib.ptr(wdev), _lib.ptr(wp), N, Cin, k, N, cin_pad, int(bf16), None))
    xd = torch.zeros(B, S, cin_pad, dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
    xd[:, :, :Cin] = x.to(dev)
    ldo = N + n_offset + extra_cols
    out = torch.full((B, S, ldo), 7.0, dtype=torch.float32, device=dev)
    out_t = torch.empty(B, N, S, dtype=torch.float32, device=dev) if want_t else None
    d = _lib.ConvDesc()
    d.B, d.S, d.Cin, d.N, d.ktaps, d.pad_left = B, S, cin_pad, N, k, k // 2
    d.lda, d.ldo, d.n_offset, d.relu = cin_pad, ldo, n_offset, int(relu)
    keep = []
    for name, t in (('bias', bias), ('scale', scale), ('shift', shift)):
        if t is not None:
            td = t.to(dev).contiguous()
            keep.append(td)
            setattr(d, name, td.data_ptr())
    if res is not None:
        rd = res.to(dev).to(torch.bfloat16 if bf16 else torch.float32).contiguous()
        keep.append(rd)
        setattr(d, 'residual_bf16' if bf16 else 'residual_f32', rd.data_ptr())
        d.ldr = N
    d.out_scale = out_scale
    d.out_f32 = out.data_ptr()
    if want_t:
        d.out_t = out_t.data_ptr()
    fn = lib.ftb_conv_gemm_bf16 if bf16 else lib.ftb_conv_gemm_f32
    _lib.check(fn(_lib.ptr(xd), _lib.ptr(wp), C.byref(d), _lib.current_stream(dev)))
    torch.cuda.synchronize()
    return out.cpu(), (out_t.cpu() if want_t else None)


CASES = [  # B, S, Cin, N, k, relu, bn, bias, res
    (2, 37, 64, 256, 5, True, True, False, False),     # predictor conv0
    (3, 50, 256, 256, 16, True, True, False, False),   # widest (even) bank kernel
    (2, 129, 256, 256, 2, True, True, False, False),   # even kernel, tile boundary at 128
    (2, 200, 80, 256, 7, True, True, False, False),    # postnet bank, Cin padded 80 -> 128
    (2, 64, 256, 80, 3, False, True, False, True),     # conv_project2: BN only + residual, N = 80
    (4, 33, 256, 512, 1, False, False, True, False),   # highway W1|W2
    (1, 300, 512, 384, 1, False, False, True, False),  # RNN input projection
    (2, 5, 128, 128, 9, True, False, True, False),     # FastPitch conv1 (bias then ReLU), S < taps
]


@pytest.mark.parametrize('B,S,Cin,N,k,relu,bn,bias,res', CASES)
@pytest.mark.parametrize('bf16', [False, True])
def test_conv_gemm(B, S, Cin, N, k, relu, bn, bias, res, bf16):
    g = torch.Generator().manual_seed(B * 7 + S + k)
    x = torch.randn(B, S, Cin, generator=g)
    w = torch.randn(N, Cin, k, generator=g) / (Cin * k) ** 0.5
    scale = torch.rand(N, generator=g) + 0.5 if bn else None
    shift = torch.randn(N, generator=g) * 0.1 if bn else None
    bvec = torch.randn(N, generator=g) * 0.1 if bias else None
    r = torch.randn(B, S, N, generator=g) if res else None
    if bf16:  # compare against the same bf16-rounded operands: isolates the kernel from input rounding
        x, w = x.bfloat16().float(), w.bfloat16().float()
        r = r.bfloat16().float() if res else None
    want = reference(x, w, k, relu, scale, shift, bvec, r, 0.5)
    got, got_t = run_kernel(x, w, k, relu, scale, shift, bvec, r, 0.5, bf16, want_t=True, n_offset=16, extra_cols=8)
    tol = 2e-3 if bf16 else 2e-4
    assert float((got[:, :, 16:16 + N] - want).abs().max()) < tol
    assert float((got_t - want.transpose(1, 2)).abs().max()) < tol
    # columns outside [n_offset, n_offset+N) are never written
    assert torch.all(got[:, :, :16] == 7.0) and torch.all(got[:, :, 16 + N:] == 7.0)
    assert _lib.lib().ftb_tc_timeout_count() == 0


@pytest.mark.parametrize('res', [False, True])
def test_tc_cta_pair_mode(res):
    """A launch big enough for the CTA-pair path (tcgen05 cta_group::2: two row tiles per MMA, the weight tile split over
    the two CTAs): 7 x 21 = 147 row tiles -- an ODD count, so the last pair has a row tile that does not exist (its loads
    zero-fill, its stores clip, its residual reads are skipped) -- times two 256-wide column tiles = 148 pair tiles."""
    g = torch.Generator().manual_seed(11)
    B, S, Cin, N, k = 7, 2600, 128, 512, 3
    x = (torch.randn(B, S, Cin, generator=g) * 0.5).bfloat16().float()
    w = (torch.randn(N, Cin, k, generator=g) / (Cin * k) ** 0.5).bfloat16().float()
    bvec = torch.randn(N, generator=g) * 0.1
    r = torch.randn(B, S, N, generator=g).bfloat16().float() if res else None
    want = reference(x, w, k, True, None, None, bvec, r, 1.0)
    got, got_t = run_kernel(x, w, k, True, None, None, bvec, r, 1.0, True, want_t=True)
    assert float((got - want).abs().max()) < 3e-3
    assert float((got_t - want.transpose(1, 2)).abs().max()) < 3e-3
    assert _lib.lib().ftb_tc_timeout_count() == 0


def test_tc_large_k_and_batch():
    """conv_project1-like: K = 3*4096, several M and N tiles."""
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(3, 260, 4096, generator=g) * 0.5).bfloat16().float()
    w = (torch.randn(256, 4096, 3, generator=g) / 110.0).bfloat16().float()
    want = reference(x, w, 3, True, None, None, None, None, 1.0)
    got, _ = run_kernel(x, w, 3, True, None, None, None, None, 1.0, True)
    assert float((got - want).abs().max()) < 5e-3
    assert _lib.lib().ftb_tc_timeout_count() == 0


@pytest.mark.parametrize('B,S,Cin,ch,K', [(2, 50, 256, 256, 16), (3, 300, 80, 256, 8), (1, 127, 64, 256, 3),
                                          (2, 128, 64, 128, 4), (1, 255, 128, 256, 5)])
@pytest.mark.parametrize('pool', [True, False])
def test_conv_bank_grouped(B, S, Cin, ch, K, pool):
    """The whole CBHG bank (k = 1..K, ReLU -> BN, concat) + MaxPool1d(2,1,1)[:S] in one grouped launch."""
    lib = _lib.lib()
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(S + K)
    x = torch.randn(B, S, Cin, generator=g).bfloat16().float()
    cin_pad = (Cin + 63) // 64 * 64
    xd = torch.zeros(B, S, cin_pad, dtype=torch.bfloat16, device=dev)
    xd[:, :, :Cin] = x.to(dev)
    out = torch.full((B, S, K * ch + 8), 7.0, dtype=torch.float32, device=dev)
    descs = (_lib.ConvDesc * K)()
    wptrs = (C.c_void_p * K)()
    keep, wants = [], []
    for i in range(K):
        k = i + 1
        w = (torch.randn(ch, Cin, k, generator=g) / (Cin * k) ** 0.5).bfloat16().float()
        scale = torch.rand(ch, generator=g) - 0.3          # negative scales: BN does not commute with the pool
        shift = torch.randn(ch, generator=g) * 0.1
        wants.append(reference(x, w, k, True, scale, shift, None, None, 1.0))
        wp = torch.empty(ch * k * cin_pad, dtype=torch.bfloat16, device=dev)
        wdev = w.to(dev).contiguous()
        _lib.check(lib.ftb_pack_conv_weight(_lib.ptr(wdev), _lib.ptr(wp), ch, Cin, k, ch, cin_pad, 1, None))
        sd, hd = scale.to(dev), shift.to(dev)
        keep += [wp, wdev, sd, hd]
        d = descs[i]
        d.B, d.S, d.Cin, d.N, d.ktaps, d.pad_left = B, S, cin_pad, ch, k, k // 2
        d.lda, d.ldo, d.n_offset, d.relu = cin_pad, K * ch + 8, i * ch, 1
        d.scale, d.shift, d.out_scale, d.out_f32 = sd.data_ptr(), hd.data_ptr(), 1.0, out.data_ptr()
        wptrs[i] = wp.data_ptr()
    _lib.check(lib.ftb_conv_bank_bf16(_lib.ptr(xd), wptrs, descs, K, int(pool), _lib.current_stream(dev)))
    torch.cuda.synchronize()
    want = torch.cat(wants, dim=2)
    if pool:
        want = F.max_pool1d(want.transpose(1, 2), kernel_size=2, stride=1, padding=1)[:, :, :S].transpose(1, 2)
    got = out.cpu()
    assert float((got[:, :, :K * ch] - want).abs().max()) < 3e-3
    assert torch.all(got[:, :, K * ch:] == 7.0)
    assert lib.ftb_tc_timeout_count() == 0


@pytest.mark.parametrize('fp16', [False, True])
@pytest.mark.parametrize('B,S,Cin,N', [(2, 150, 1024, 80), (3, 70, 512, 80), (1, 129, 64, 256)])
def test_linear_over_hi_lo_pair(B, S, Cin, N, fp16):
    """The output heads: activation and weight both as 16-bit pairs hi + lo, three part products on tcgen05.  Against an
    fp64 product of the fp32 operands: the error is that of a 16-bit (bf16 pair) / 21-bit (half pair) operand, orders
    below the single-part kernel's."""
    lib, dev = _lib.lib(), torch.device('cuda')
    g = torch.Generator().manual_seed(Cin + N + S)
    x = torch.randn(B, S, Cin, generator=g) * 0.3
    w = torch.randn(N, Cin, generator=g) * (30.0 / Cin ** 0.5)            # trained-magnitude head
    bias = torch.randn(N, generator=g)
    dt = torch.float16 if fp16 else torch.bfloat16
    hi = x.to(dt)
    lo = (x - hi.float()).to(dt)
    pair = torch.cat([hi, lo], dim=2).contiguous().to(dev)
    wd = w.to(dev).contiguous()
    wp = torch.empty(N * 3 * Cin, dtype=dt, device=dev)
    _lib.check(lib.ftb_pack_conv_weight(_lib.ptr(wd), _lib.ptr(wp), N, Cin, 1, N, Cin, 5 if fp16 else 4, None))
    bd = bias.to(dev)
    out = torch.full((B, S, N + 4), 7.0, dtype=torch.float32, device=dev)
    _lib.check(lib.ftb_linear_pair(_lib.ptr(pair), _lib.ptr(wp), B, S, Cin, N, _lib.ptr(bd), _lib.ptr(out), N + 4, int(fp16),
                                   _lib.current_stream(dev)))
    torch.cuda.synchronize()
    want = (x.double() @ w.double().T + bias.double()).float()
    got = out.cpu()
    single = (hi.float().double() @ w.to(dt).float().double().T + bias.double()).float()  # what one 16-bit part gives
    e_pair, e_single = float((got[:, :, :N] - want).abs().max()), float((single - want).abs().max())
    print(f'fp16={fp16} Cin={Cin}: pair max-abs {e_pair:.2e}, single-part {e_single:.2e}, |y| max {float(want.abs().max()):.1f}')
    assert e_pair < (2e-5 if fp16 else 6e-4) * max(1.0, float(want.abs().max())) and e_pair < e_single / 20
    assert torch.all(got[:, :, N:] == 7.0) and lib.ftb_tc_timeout_count() == 0
