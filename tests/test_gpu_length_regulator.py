"""LengthRegulator / duration fallback on the GPU: bit-exact against the reference fixture and the oracle."""
import ctypes as C

import pytest
import torch

from forwardtacotron_b200 import _lib
from forwardtacotron_b200.models.common_layers import LengthRegulator
from oracle import model_oracle as mo

from util import load

pytestmark = pytest.mark.gpu


def test_reference_fixture_bit_exact():
    g = load('length_regulator')
    dur = g['dur_in'].clone().cuda()
    y = LengthRegulator()(g['x'].cuda(), dur)
    assert torch.equal(y.cpu(), g['y'])
    assert torch.equal(dur.cpu(), g['dur_out'])          # clamped in place like the reference


@pytest.mark.parametrize('B,T,Cn,dtype', [(3, 50, 512, torch.float32), (64, 200, 512, torch.bfloat16),
                                          (2, 2000, 256, torch.bfloat16), (1, 1, 8, torch.float32),
                                          (5, 777, 1024, torch.float32)])
def test_random_against_oracle(B, T, Cn, dtype):
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, Cn, generator=g).to(dtype)
    dur = torch.rand(B, T, generator=g) * 12 - 1.5
    dur[0, : min(T, 4)] = torch.tensor([0.5, 1.4999999, 2.5, -3.0])[: min(T, 4)]
    d_or = dur.clone()
    want = mo.length_regulate(x.float(), d_or).to(dtype)
    d_gpu = dur.clone().cuda()
    got = LengthRegulator()(x.cuda(), d_gpu)
    assert got.dtype == dtype and torch.equal(got.cpu(), want)
    assert torch.equal(d_gpu.cpu(), d_or)


def test_plan_outputs_are_exact_integers():
    g = torch.Generator().manual_seed(9)
    dur = (torch.rand(7, 333, generator=g) * 9 - 1).cuda()
    want = (dur.cpu().clamp(min=0) + 0.5).long()
    cum, total = LengthRegulator.plan(dur)
    assert torch.equal(cum.cpu().long(), want.cumsum(1))
    assert torch.equal(total.cpu().long(), want.sum(1))


def test_all_zero_row_and_truncating_L():
    x = torch.arange(2 * 3 * 8, dtype=torch.float32).view(2, 3, 8).cuda()
    dur = torch.tensor([[0.0, 0.2, 0.4], [1.0, 2.0, 0.0]]).cuda()
    y = LengthRegulator()(x, dur)
    assert y.shape == (2, 3, 8) and float(y[0].abs().sum()) == 0.0
    assert torch.equal(y[1, 0], x[1, 0]) and torch.equal(y[1, 1], x[1, 1]) and torch.equal(y[1, 2], x[1, 1])


@pytest.mark.parametrize('case', ['negative_sum', 'zero', 'positive', 'mixed_trunc'])
def test_duration_fallback(case):
    dur = {'negative_sum': torch.tensor([[-1.5, 0.9, 0.2], [0.1, -2.0, 0.99]]),
           'zero': torch.zeros(2, 3),
           'positive': torch.tensor([[0.2, 1.0, 0.3], [0.0, 0.0, 0.0]]),
           # trunc toward zero: -0.9 -> 0, 0.9 -> 0  => sum 0 => fallback
           'mixed_trunc': torch.tensor([[-0.9, 0.9, 0.5], [0.99, -0.99, 0.0]])}[case]
    want = mo.apply_duration_fallback(dur.clone())
    d = dur.clone().cuda()
    scratch = torch.zeros(8, dtype=torch.uint8, device='cuda')
    _lib.check(_lib.lib().ftb_duration_fallback(_lib.ptr(d), d.numel(), _lib.ptr(scratch), _lib.current_stream(d.device)))
    assert torch.equal(d.cpu(), want)


@pytest.mark.parametrize('B,T,L_cut', [(3, 50, 0), (64, 200, 0), (2, 2000, 0), (4, 33, 17), (1, 1, 0)])
def test_frame_index_matches_the_expansion(B, T, L_cut):
    """ftb_length_index is the LengthRegulator as a gather map: x_rows[idx] == expand(x), pad row for the zero tail."""
    g = torch.Generator().manual_seed(B * 31 + T)
    dur = (torch.rand(B, T, generator=g) * 9 - 1).cuda()
    cum, total = LengthRegulator.plan(dur)
    L = int(total.max()) if not L_cut else min(L_cut, int(total.max()))
    if L == 0:
        pytest.skip('all durations rounded to zero')
    idx = torch.empty(B, L, dtype=torch.int32, device='cuda')
    _lib.check(_lib.lib().ftb_length_index(_lib.ptr(cum), _lib.ptr(idx), B, T, L, B * T, _lib.current_stream(idx.device)))
    x = torch.randn(B, T, 64, generator=g).cuda()
    rows = torch.cat([x.view(B * T, 64), torch.zeros(1, 64, device='cuda')])
    want = LengthRegulator.expand(x, cum, L)
    assert torch.equal(rows[idx.long()], want)
    # and against the oracle's own repeat_interleave
    r = (dur.cpu().clamp(min=0) + 0.5).long()
    for b in range(B):
        seq = torch.repeat_interleave(torch.arange(T) + b * T, r[b])[:L]
        assert torch.equal(idx[b, :len(seq)].cpu().long(), seq) and bool((idx[b, len(seq):] == B * T).all())
