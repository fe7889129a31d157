"""DSP oracle: the numpy restatement of librosa's stft + Slaney mel against torchaudio's independent
implementation, the frozen fixture, and the facts the reference's own golden pins (tests/test_dsp.py)."""
import numpy as np
import pytest
import torch

from oracle import dsp_oracle

from util import GOLD


def test_against_torchaudio():
    torchaudio = pytest.importorskip('torchaudio')
    rng = np.random.default_rng(5)
    y = (0.1 * rng.standard_normal(22050)).astype(np.float32)
    mel = dsp_oracle.wav_to_mel(y)
    ta = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, f_min=0,
                                              f_max=8000, n_mels=80, power=1.0, center=True, pad_mode='reflect',
                                              norm='slaney', mel_scale='slaney')
    ref = torch.log(torch.clamp(ta(torch.from_numpy(y)), min=1e-5)).numpy()
    assert mel.shape == ref.shape == (80, 1 + 22050 // 256)
    assert np.abs(mel - ref).max() < 1e-4
    assert np.abs(dsp_oracle.mel_filterbank(22050, 1024, 80, 0, 8000) - ta.mel_scale.fb.numpy().T).max() < 1e-6


def test_filterbank_structure():
    fb = dsp_oracle.mel_filterbank(22050, 1024, 80, 0, 8000)
    assert fb.shape == (80, 513) and fb.dtype == np.float32
    assert int((fb != 0).sum()) == 727          # SURVEY 8c [probe]
    assert int(np.nonzero(fb.any(0))[0].max()) == 371


def test_frozen_fixture_and_reference_golden_facts():
    g = np.load(GOLD / 'dsp_noise10k.npz')
    mel = dsp_oracle.wav_to_mel(g['y'])
    assert np.abs(mel - g['mel']).max() < 1e-5
    # tests/test_dsp.py:18-25 pins (80, 40) float32 for 10 000 samples, floor log(1e-5)
    assert mel.shape == (80, 40) and mel.dtype == np.float32
    silent = dsp_oracle.wav_to_mel(np.zeros(10000, np.float32))
    assert np.allclose(silent, np.log(1e-5))
    assert np.isclose(silent.min(), -11.512925)


def test_unnormalized_and_short_clip():
    rng = np.random.default_rng(1)
    y = rng.standard_normal(700).astype(np.float32)
    lin = dsp_oracle.wav_to_mel(y, normalize=False)
    assert lin.shape == (80, 1 + 700 // 256) and (lin >= 0).all()
    assert np.allclose(np.log(np.clip(lin, 1e-5, None)), dsp_oracle.wav_to_mel(y), atol=1e-6)


def test_inverse_path_restatements_are_self_consistent():
    """istft(stft(y)) == y, stft / istft agree with torch's implementations, NNLS fits its target, trim bounds hold."""
    import torch
    from oracle import dsp_oracle as d
    rng = np.random.default_rng(0)
    y = (0.1 * rng.standard_normal(22050)).astype(np.float32)
    S = d.stft(y)
    yr = d.istft(S)
    assert len(yr) == 256 * (S.shape[1] - 1) and np.abs(yr - y[:len(yr)]).max() < 1e-6
    win = torch.hann_window(1024, periodic=True)
    St = torch.stft(torch.from_numpy(y), 1024, 256, 1024, window=win, center=True, pad_mode='reflect', return_complex=True)
    assert np.abs(St.numpy() - S).max() < 1e-5
    yt = torch.istft(St, 1024, 256, 1024, window=win, center=True).numpy()
    assert np.abs(yt - yr[:len(yt)]).max() < 1e-6
    M = np.exp(d.wav_to_mel(y)).astype(np.float32)
    X = d.mel_to_stft(M)
    A = d.mel_filterbank(22050, 1024, 80, 0, 8000)
    assert X.min() >= 0 and np.abs(A @ X - M).max() < 1e-2 * M.max()
    yy = np.concatenate([np.zeros(5000, np.float32), y, np.zeros(7000, np.float32)])
    cut, (s, e) = d.trim_silence(yy, 60)
    assert (s, e) == (4096, 28160) and len(cut) == e - s
    w = d.griffinlim(np.abs(S), np.exp(2j * np.pi * rng.random(S.shape)), n_iter=4)
    assert w.shape == yr.shape and np.isfinite(w).all()
