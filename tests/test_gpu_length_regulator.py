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
