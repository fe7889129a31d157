"""Pins the oracle against the REAL reference and freezes golden fixtures.

Runs only in the build container (it imports /root/reference, which does not travel to the GPU
box).  For ForwardTacotron and FastPitch it
  1. builds synthetic weights with forwardtacotron_b200.utils.synth (our mirror's state_dict),
  2. loads them with strict=True into the reference's own classes (proves the layout contract),
  3. runs the reference's generate() and oracle.model_oracle on the same inputs and asserts they
     agree (durations exactly, floats to 1e-5),
  4. stores small input/output fixtures under tests/golden/ for the CPU and GPU test-suites.

    python oracle/make_golden.py [--calibrate]
"""
from __future__ import annotations

import argparse
import copy
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path('/root/reference')

from forwardtacotron_b200.utils import synth  # noqa: E402
from forwardtacotron_b200.utils.config import default_config  # noqa: E402
from oracle import dsp_oracle, model_oracle as mo  # noqa: E402

GOLD = ROOT / 'tests' / 'golden'


def import_reference():
    if not REF.exists():
        raise SystemExit('/root/reference is not present: golden generation runs in the build container only')
    sys.path.insert(0, str(REF))
    from models.fast_pitch import FastPitch as RefFP  # type: ignore
    from models.forward_tacotron import ForwardTacotron as RefFT  # type: ignore
    from models.common_layers import LengthRegulator as RefLR  # type: ignore
    return RefFT, RefFP, RefLR


def max_abs(a, b):
    return float((a - b).abs().max()) if a.numel() else 0.0


def case(model_type: str, RefCls, B: int, T: int, *, ragged=False, alpha=1.0, plain=False, seed=0, callbacks=False,
         mel_gain=1.0):
    model, cfg = synth.synthetic_model(model_type, seed=seed, plain_init=plain, mel_gain=mel_gain)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref = RefCls.from_config(copy.deepcopy(cfg))
    ref.load_state_dict(sd, strict=True)  # the layout contract
    x = synth.synthetic_tokens(B, T, seed=1 + seed, ragged=ragged)
    pf = (lambda p: p * 1.2) if callbacks else (lambda p: p)
    ef = (lambda e: e + 0.1) if callbacks else (lambda e: e)
    with torch.no_grad():
        r = ref.generate(x, alpha=alpha, pitch_function=pf, energy_function=ef)
        gen = mo.ft_generate if model_type == 'forward_tacotron' else mo.fp_generate
        o = gen(sd, x, alpha=alpha, pitch_function=pf, energy_function=ef)
    reps_r = (r['dur'] + 0.5).long()
    reps_o = (o['dur'] + 0.5).long()
    assert torch.equal(reps_r, reps_o), f'{model_type}: oracle durations differ from the reference'
    errs = {k: max_abs(r[k], o[k]) for k in ('mel', 'mel_post', 'dur', 'pitch', 'energy')}
    assert max(errs.values()) < 2e-5, f'{model_type}: oracle deviates from the reference: {errs}'
    return sd, cfg, x, r, errs


def save_case(name, x, r, extra=None):
    out = {'x': x.numpy(), 'dur': r['dur'].numpy(), 'pitch': r['pitch'].numpy(), 'energy': r['energy'].numpy(),
           'mel': r['mel'].numpy(), 'mel_post': r['mel_post'].numpy()}
    out.update(extra or {})
    np.savez_compressed(GOLD / f'{name}.npz', **out)


def calibrate():
    for mt, gen in (('forward_tacotron', mo.ft_predict), ('fast_pitch', mo.fp_predict)):
        for s, b in ((None, None), (60, 6), (3, 6)):
            model, _ = synth.synthetic_model(mt, dur_scale=s, dur_bias=b)
            x = synth.synthetic_tokens(16, 200)
            with torch.no_grad():
                dur = gen(model.state_dict(), x)[0]
            r = (dur.clamp(min=0) + 0.5).long().float()
            print(f'{mt} s={s} b={b}: dur mean {r.mean():.2f} std {r.std():.2f} min {r.min():.0f} max {r.max():.0f}')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--calibrate', action='store_true')
    args = ap.parse_args()
    torch.set_num_threads(8)
    if args.calibrate:
        calibrate()
        return
    RefFT, RefFP, RefLR = import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)

    # --- LengthRegulator: bit-exact incl. negatives, x.5 ties, zeros, ragged totals
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(5, 37, 16, generator=g)
    dur = torch.rand(5, 37, generator=g) * 5 - 0.7
    dur[0, :6] = torch.tensor([0.5, 1.5, 2.5, -0.5, 0.49999997, 0.0])
    dur[3] = 0.0  # an utterance with no frames at all
    d_ref, d_or = dur.clone(), dur.clone()
    y_ref = RefLR()(xs, d_ref)
    y_or = mo.length_regulate(xs, d_or)
    assert torch.equal(y_ref, y_or) and torch.equal(d_ref, d_or)
    np.savez_compressed(GOLD / 'length_regulator.npz', x=xs.numpy(), dur_in=dur.numpy(), dur_out=d_ref.numpy(),
                        y=y_ref.numpy())
    print('length_regulator ok', tuple(y_ref.shape))

    # --- explicit GRU/LSTM restatement == ATen
    model, _ = synth.synthetic_model('forward_tacotron')
    sd = model.state_dict()
    xr = torch.randn(2, 9, 256, generator=g)
    assert max_abs(mo.rnn_explicit(sd, 'prenet.rnn', xr, 'gru'), mo.rnn(sd, 'prenet.rnn', xr, 'gru')) < 1e-5
    xr = torch.randn(2, 7, 512, generator=g)
    assert max_abs(mo.rnn_explicit(sd, 'lstm', xr, 'lstm'), mo.rnn(sd, 'lstm', xr, 'lstm')) < 1e-5
    print('rnn_explicit ok')

    # --- ForwardTacotron
    sd, cfg, x, r, e = case('forward_tacotron', RefFT, 2, 24)
    save_case('ft_b2_t24', x, r)
    print('ft_b2_t24', e, 'L', r['mel'].shape[-1])
    sd, cfg, x, r, e = case('forward_tacotron', RefFT, 3, 40, ragged=True, alpha=1.1, callbacks=True)
    save_case('ft_b3_t40_ragged', x, r, {'alpha': np.float32(1.1)})
    print('ft_b3_t40_ragged', e, 'L', r['mel'].shape[-1])
    sd, cfg, x, r, e = case('forward_tacotron', RefFT, 2, 16, plain=True)
    assert float(r['dur'].min()) == 2.0 and float(r['dur'].max()) == 2.0  # fallback branch
    save_case('ft_b2_t16_fallback', x, r)
    print('ft_b2_t16_fallback', e)
    # CBHG / SeriesPredictor sub-module fixtures (rows a3/a4)
    model, _ = synth.synthetic_model('forward_tacotron')
    ref = RefFT.from_config(copy.deepcopy(cfg))
    ref.load_state_dict(model.state_dict(), strict=True)
    ref.eval()
    with torch.no_grad():
        xin = torch.randn(2, 256, 33, generator=g)
        y_pre = ref.prenet(xin)
        assert max_abs(y_pre, mo.cbhg(model.state_dict(), 'prenet', xin)) < 2e-5
        min_ = torch.randn(2, 80, 50, generator=g) * 0.5
        y_post = ref.postnet(min_)
        assert max_abs(y_post, mo.cbhg(model.state_dict(), 'postnet', min_)) < 2e-5
        xt = synth.synthetic_tokens(2, 30, seed=5)
        y_dur = ref.dur_pred(xt, alpha=0.9)
        assert max_abs(y_dur, mo.ft_series_predictor(model.state_dict(), 'dur_pred', xt, 0.9)) < 2e-5
    np.savez_compressed(GOLD / 'ft_submodules.npz', prenet_in=xin.numpy(), prenet_out=y_pre.numpy(),
                        postnet_in=min_.numpy(), postnet_out=y_post.numpy(), dur_tokens=xt.numpy(),
                        dur_out=y_dur.numpy())
    print('ft_submodules ok')

    # --- FastPitch
    sd, cfg, x, r, e = case('fast_pitch', RefFP, 2, 24, callbacks=True)
    save_case('fp_b2_t24', x, r)
    print('fp_b2_t24', e, 'L', r['mel'].shape[-1])
    sd, cfg, x, r, e = case('fast_pitch', RefFP, 3, 40, ragged=True, alpha=0.9)
    save_case('fp_b3_t40_ragged', x, r, {'alpha': np.float32(0.9)})
    print('fp_b3_t40_ragged', e, 'L', r['mel'].shape[-1])

    # --- DSP: numpy restatement vs torchaudio's independent Slaney implementation
    import torchaudio
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal(10000)).astype(np.float32)
    mel = dsp_oracle.wav_to_mel(y)
    assert mel.shape == (80, 1 + 10000 // 256) and mel.dtype == np.float32
    ta = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, f_min=0,
                                              f_max=8000, n_mels=80, power=1.0, center=True, pad_mode='reflect',
                                              norm='slaney', mel_scale='slaney')
    mel_ta = torch.log(torch.clamp(ta(torch.from_numpy(y)), min=1e-5)).numpy()
    err = float(np.abs(mel - mel_ta).max())
    assert err < 1e-4, err
    fb_err = float(np.abs(dsp_oracle.mel_filterbank(22050, 1024, 80, 0, 8000) - ta.mel_scale.fb.numpy().T).max())
    assert fb_err < 1e-6, fb_err
    ref_fixture = np.load(REF / 'tests' / 'resources' / 'test_mel.npy')
    assert ref_fixture.shape == mel.shape and ref_fixture.dtype == mel.dtype
    assert np.isclose(ref_fixture.min(), np.log(1e-5))
    np.savez_compressed(GOLD / 'dsp_noise10k.npz', y=y, mel=mel)
    print(f'dsp ok: vs torchaudio {err:.2e}, filterbank {fb_err:.2e}; reference fixture shape/dtype/floor match')


if __name__ == '__main__':
    main()
