"""16-bit multi-head attention core (csrc/attention_umma.cu: tcgen05 / TMEM; csrc/attention.cu: mma.sync) against a plain
fp32 torch evaluation of nn.MultiheadAttention's core (models/fast_pitch.py:64,80-82) on the same 16-bit q / k / v."""
import pytest
import torch

from forwardtacotron_b200 import _lib

pytestmark = pytest.mark.gpu


def reference(qkv, tokens, heads):
    B, S, E3 = qkv.shape
    E = E3 // 3
    hd = E // heads
    q, k, v = (t.float().view(B, S, heads, hd).transpose(1, 2) for t in qkv.split(E, dim=-1))
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    if tokens is not None:
        s = s.masked_fill((tokens == 0)[:, None, None, :], float('-inf'))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)          # a row whose keys are all masked: the kernels write zeros
    return (p @ v).transpose(1, 2).reshape(B, S, E)


def run(qkv, tokens, heads, impl):
    B, S, E3 = qkv.shape
    E = E3 // 3
    ctx = torch.full((B, S, E), float('nan'), dtype=qkv.dtype, device=qkv.device)
    _lib.check(_lib.lib().ftb_attention_16(_lib.ptr(qkv), _lib.ptr(tokens) if tokens is not None else None, _lib.ptr(ctx),
                                           B, S, E, heads, int(qkv.dtype == torch.float16), impl,
                                           _lib.current_stream(qkv.device)))
    torch.cuda.synchronize()
    return ctx


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('B,S,E,heads,masked', [
    (2, 300, 256, 2, True),      # FastPitch prenet shape: hd 128, padded tokens
    (3, 1000, 256, 2, False),    # postnet-like: several key tiles, no mask
    (2, 129, 256, 2, True),      # one row past a 128-key tile
    (1, 5, 256, 2, False),       # shorter than one tile
    (4, 333, 128, 2, True),      # pitch / energy predictor: hd 64
    (1, 2047, 256, 2, True),     # cfg3 / cfg5 length
])
def test_attention_matches_fp32_reference(impl, dtype, B, S, E, heads, masked):
    g = torch.Generator().manual_seed(S * 7 + E)
    qkv = torch.randn(B, S, 3 * E, generator=g).mul(1.5).to(dtype).cuda()
    tokens = None
    if masked:
        tokens = torch.randint(1, 50, (B, S), generator=g)
        for b in range(B):
            tokens[b, S - (b * 37) % max(1, S // 2):] = 0 if b else tokens[b, -1]
        tokens[:, S // 3] = 0                 # a masked key in the middle as well
        tokens = tokens.cuda()
    want = reference(qkv, tokens, heads)
    got = run(qkv, tokens, heads, impl).float()
    assert torch.isfinite(got).all()
    d = (got - want).abs()
    tol = 4e-3 if dtype == torch.float16 else 3e-2     # 16-bit P and output rounding; |ctx| ~ 1
    assert float(d.max()) < tol and float(d.mean()) < tol / 8, (float(d.max()), float(d.mean()))
    assert _lib.lib().ftb_tc_timeout_count() == 0


def test_sharp_rows_exercise_the_lazy_rescale():
    """Scores whose row maximum jumps by far more than 2^8 between key tiles (the tcgen05 kernel then rescales its O rows
    in tensor memory) and rows that stay flat; both kernels must agree with the reference."""
    g = torch.Generator().manual_seed(5)
    B, S, E, heads = 2, 640, 256, 2
    qkv = torch.randn(B, S, 3 * E, generator=g)
    qkv[:, :, :E] *= 4.0                                # large queries -> peaked softmax
    qkv[:, 400:, E:2 * E] *= 6.0                         # late keys dominate: the maximum moves in tile 3
    qkv = qkv.half().cuda()
    want = reference(qkv, None, heads)
    for impl in (0, 1):
        got = run(qkv, None, heads, impl).float()
        d = (got - want).abs()
        assert float(d.max()) < 2e-2 and float(d.mean()) < 1e-3, (impl, float(d.max()), float(d.mean()))
