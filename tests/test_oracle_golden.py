"""The oracle restatement reproduces the fixtures that oracle/make_golden.py froze from the REAL
reference (imported from /root/reference in the build container)."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from forwardtacotron_b200.utils import synth

from util import load


def test_length_regulator_fixture():
    g = load('length_regulator')
    dur = g['dur_in'].clone()
    y = mo.length_regulate(g['x'], dur)
    assert torch.equal(y, g['y']) and torch.equal(dur, g['dur_out'])
    assert y.shape[1] == int((g['dur_out'] + 0.5).long().sum(1).max())
    assert float(y[3].abs().sum()) == 0.0  # the all-zero-duration utterance


@pytest.mark.parametrize('name,kind,alpha,cb,ragged,plain', [
    ('ft_b2_t24', 'forward_tacotron', 1.0, False, False, False),
    ('ft_b3_t40_ragged', 'forward_tacotron', 1.1, True, True, False),
    ('ft_b2_t16_fallback', 'forward_tacotron', 1.0, False, False, True),
    ('fp_b2_t24', 'fast_pitch', 1.0, True, False, False),
    ('fp_b3_t40_ragged', 'fast_pitch', 0.9, False, True, False),
])
def test_generate_fixture(name, kind, alpha, cb, ragged, plain):
    g = load(name)
    model, _ = synth.synthetic_model(kind, plain_init=plain)
    sd = model.state_dict()
    pf = (lambda p: p * 1.2) if cb else (lambda p: p)
    ef = (lambda e: e + 0.1) if cb else (lambda e: e)
    gen = mo.ft_generate if kind == 'forward_tacotron' else mo.fp_generate
    o = gen(sd, g['x'], alpha=alpha, pitch_function=pf, energy_function=ef)
    assert torch.equal((o['dur'] + 0.5).long(), (g['dur'] + 0.5).long())
    for k in ('mel', 'mel_post', 'dur', 'pitch', 'energy'):
        assert o[k].shape == g[k].shape, k
        assert float((o[k] - g[k]).abs().max()) < 5e-5, k
    if plain:
        assert float(o['dur'].min()) == 2.0 == float(o['dur'].max())


def test_submodule_fixtures():
    g = load('ft_submodules')
    model, _ = synth.synthetic_model('forward_tacotron')
    sd = model.state_dict()
    assert float((mo.cbhg(sd, 'prenet', g['prenet_in']) - g['prenet_out']).abs().max()) < 5e-5
    assert float((mo.cbhg(sd, 'postnet', g['postnet_in']) - g['postnet_out']).abs().max()) < 5e-5
    assert float((mo.ft_series_predictor(sd, 'dur_pred', g['dur_tokens'], 0.9) - g['dur_out']).abs().max()) < 5e-5


def test_explicit_rnn_equations_match_aten():
    model, _ = synth.synthetic_model('forward_tacotron')
    sd = model.state_dict()
    x = torch.randn(2, 6, 256, generator=torch.Generator().manual_seed(0))
    assert float((mo.rnn_explicit(sd, 'postnet.rnn', x, 'gru') - mo.rnn(sd, 'postnet.rnn', x, 'gru')).abs().max()) < 1e-5
    x = torch.randn(2, 5, 512, generator=torch.Generator().manual_seed(0))
    assert float((mo.rnn_explicit(sd, 'lstm', x, 'lstm') - mo.rnn(sd, 'lstm', x, 'lstm')).abs().max()) < 1e-5


def test_even_kernel_taps():
    """k even: left pad k//2, right pad k-1-k//2 after truncation (SURVEY 7 'even-width convs')."""
    g = torch.Generator().manual_seed(1)
    for k in (2, 4, 16):
        w = torch.randn(3, 5, k, generator=g)
        x = torch.randn(2, 5, 11, generator=g)
        ref = torch.nn.functional.conv1d(x, w, padding=k // 2)[:, :, :11]
        xp = torch.nn.functional.pad(x, (k // 2, k - 1 - k // 2))
        man = torch.stack([sum(w[:, :, j] @ xp[b, :, j:j + 11] for j in range(k)) for b in range(2)])
        assert float((ref - man).abs().max()) < 1e-5


def test_fastpitch_max_len_raises():
    model, _ = synth.synthetic_model('fast_pitch')
    with pytest.raises(RuntimeError, match='must match the size'):
        mo.forward_transformer(model.state_dict(), 'prenet', torch.zeros(1, 5001, 256), 2)


def _forward_batch(g):
    return {'x': g['x'], 'dur': g['dur_in'].clone(), 'mel_len': g['mel_len'], 'pitch': g['pitch_in'],
            'energy': g['energy_in'], 'mel': torch.zeros(g['x'].shape[0], 80, int(g['mel_frames']))}


def test_forward_fixture():
    """Teacher-forced forward() in eval mode (GTA features): the oracle against the reference's own output
    (oracle/make_golden_forward.py), including the packed-sequence padding."""
    g = load('ft_forward_b3_t30')
    model, _ = synth.synthetic_model('forward_tacotron')
    out = mo.ft_forward(model.state_dict(), _forward_batch(g), model.pitch_strength, model.energy_strength)
    for k in ('mel', 'mel_post', 'dur', 'pitch', 'energy'):
        assert out[k].shape == g[k].shape and float((out[k] - g[k]).abs().max()) < 2e-5, k
    n = int(g['mel_len'].max())
    assert torch.all(out['mel'][:, :, n:] == -11.5129) and torch.all(out['mel_post'][:, :, n:] == -11.5129)


def test_fastpitch_forward_fixture():
    """FastPitch.forward in eval mode: token padding mask on predictors / prenet, mel-length key mask on the postnet."""
    g = load('fp_forward_b3_t30')
    model, _ = synth.synthetic_model('fast_pitch')
    out = mo.fp_forward(model.state_dict(), _forward_batch(g), model.pitch_strength, model.energy_strength)
    for k in ('mel', 'mel_post', 'dur', 'pitch', 'energy'):
        assert out[k].shape == g[k].shape and float((out[k] - g[k]).abs().max()) < 2e-5, k
    assert bool((g['x'] == 0).any())                                   # the fixture really has padded tokens
    assert torch.all(out['mel'][:, :, int((g['dur_in'].clamp(min=0) + 0.5).long().sum(1).max()):] == -11.5129)
