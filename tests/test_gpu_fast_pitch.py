"""FastPitch.generate on the GPU against the reference fixtures and the oracle (rows a10-a12)."""
import pytest
import torch

from oracle import model_oracle as mo

from util import MAX_ABS, MEAN_ABS, assert_close, cpu_state_dict, cuda_model, load, near_tie_mask, rounded
from forwardtacotron_b200.utils import synth

pytestmark = pytest.mark.gpu


def check_against(model, x, want, alpha=1.0, pf=None, ef=None):
    pf = pf or (lambda p: p)
    ef = ef or (lambda e: e)
    out = model.generate(x.cuda(), alpha=alpha, pitch_function=pf, energy_function=ef)
    assert_close(out['dur'], want['dur'], 1e-3, 1e-4, 'dur')
    assert_close(out['pitch'], want['pitch'], what='pitch')
    assert_close(out['energy'], want['energy'], what='energy')
    flips = rounded(out['dur']) != rounded(want['dur'])
    if flips.any():
        assert bool((near_tie_mask(want['dur']) | ~flips).all()), 'a duration differs that is not a rounding near-tie'
        out = model.synthesize(x.cuda(), want['dur'].clone().cuda(), pf(out['pitch']), ef(out['energy']))
    assert out['mel_post'] is out['mel']                       # same tensor, as in the reference (:339)
    return out, assert_close(out['mel'], want['mel'], what='mel')


@pytest.mark.parametrize('gemm_mode', [1, 0])
@pytest.mark.parametrize('name,alpha,cb', [('fp_b2_t24', 1.0, True), ('fp_b3_t40_ragged', 0.9, False)])
def test_reference_fixtures(name, alpha, cb, gemm_mode):
    g = load(name)
    model, _ = cuda_model('fast_pitch', gemm_mode)
    pf = (lambda p: p * 1.2) if cb else None
    ef = (lambda e: e + 0.1) if cb else None
    out, res = check_against(model, g['x'], g, alpha, pf, ef)
    print(name, 'gemm_mode', gemm_mode, res)


@pytest.mark.parametrize('gemm_mode', [1, 0])
def test_batch_against_oracle(gemm_mode):
    model, _ = cuda_model('fast_pitch', gemm_mode)
    x = synth.synthetic_tokens(6, 150, seed=3, ragged=True)
    want = mo.fp_generate(cpu_state_dict(model), x, pitch_function=lambda p: p * 1.2, energy_function=lambda e: e + 0.1)
    out, res = check_against(model, x, want, pf=lambda p: p * 1.2, ef=lambda e: e + 0.1)
    print('fp B6xT150 gemm_mode', gemm_mode, res, 'L', out['mel'].shape[-1])


def test_positional_table_limit_raises():
    model, _ = cuda_model('fast_pitch', 0)
    with pytest.raises(RuntimeError, match='must match the size'):
        model.generate(torch.ones(1, 5001, dtype=torch.long, device='cuda'))


def test_bf16_tensor_core_mode_is_opt_in():
    """gemm_mode 2 puts the prenet / postnet / lin GEMMs on the bf16 tcgen05 kernel.  bf16 operands miss the 1e-3
    mean-abs budget on this model, which is why it is not the default: here only a looser bound and
    exact durations are required (the duration predictor stays fp32)."""
    g = load('fp_b3_t40_ragged')
    model, _ = cuda_model('fast_pitch', 2)
    out = model.generate(g['x'].cuda(), alpha=0.9)
    assert torch.equal(rounded(out['dur']), rounded(g['dur']))
    mx, mn = assert_close(out['mel'], g['mel'], 5e-2, 5e-3, 'mel (bf16 GEMMs)')
    print('fp gemm_mode 2', mx, mn)


@pytest.mark.parametrize('gemm_mode', [1, 0])
def test_teacher_forced_forward_against_reference_fixture(gemm_mode):
    """FastPitch.forward in eval mode (models/fast_pitch.py:243-283) against the reference's own output: masked
    predictors / prenet, batch durations, postnet masked past mel_len, padded to mel.size(2)."""
    g = load('fp_forward_b3_t30')
    model, _ = cuda_model('fast_pitch', gemm_mode)
    model.eval()
    B = g['x'].shape[0]
    batch = {'x': g['x'].cuda(), 'dur': g['dur_in'].clone().cuda(), 'mel_len': g['mel_len'].cuda(),
             'pitch': g['pitch_in'].cuda(), 'energy': g['energy_in'].cuda(),
             'mel': torch.zeros(B, 80, int(g['mel_frames']), device='cuda')}
    out = model(batch)
    assert out['mel_post'] is out['mel'] and out['mel'].shape == g['mel'].shape
    tol = (1e-4, 1e-5) if gemm_mode == 1 else (MAX_ABS, MEAN_ABS)
    assert_close(out['dur'], g['dur'], 1e-4, 1e-5, 'dur_hat')
    assert_close(out['pitch'], g['pitch'], *tol, 'pitch_hat')
    assert_close(out['energy'], g['energy'], *tol, 'energy_hat')
    for b, n in enumerate(g['mel_len'].tolist()):     # the frames a GTA dump keeps
        assert_close(out['mel'][b, :, :n], g['mel'][b, :, :n], *tol, f'mel[{b}] valid frames')
    assert_close(out['mel'], g['mel'], 2 * tol[0], 2 * tol[1], 'mel (all frames)')
    model.train()
    with pytest.raises(NotImplementedError):
        model(batch)


def test_fused_layernorm_epilogue_matches_the_standalone_layernorm(monkeypatch):
    """x = norm(x + sublayer(x)) (models/fast_pitch.py:84,91) runs inside the out_proj / conv2 GEMM epilogues (d_model 256
    = one tile: the row is complete in tensor memory).  FTB_UNFUSED_LN=1 (read when a native handle is created) keeps the
    stand-alone LayerNorm launches.  Same two-pass statistics in fp32, different summation order -> equal to a few ulp of
    the 16-bit operand copy, far inside the parity budget; and both inside the budget against the oracle."""
    x = synth.synthetic_tokens(4, 130, seed=3)
    fused, _ = cuda_model('fast_pitch', 0)
    a = fused.generate(x.cuda())
    monkeypatch.setenv('FTB_UNFUSED_LN', '1')
    plain, _ = cuda_model('fast_pitch', 0)
    b = plain.generate(x.cuda())
    monkeypatch.delenv('FTB_UNFUSED_LN')
    assert torch.equal(rounded(a['dur']), rounded(b['dur']))
    d = (a['mel'] - b['mel']).abs()
    print('fused vs stand-alone LayerNorm: max-abs', float(d.max()), 'mean-abs', float(d.mean()))
    assert float(d.max()) < 5e-3 and float(d.mean()) < 5e-4
    want = mo.fp_generate(cpu_state_dict(fused), x)
    assert_close(a['mel'], want['mel'], what='fused LayerNorm vs oracle')
    from forwardtacotron_b200 import _lib
    assert _lib.lib().ftb_tc_timeout_count() == 0
