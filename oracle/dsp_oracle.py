"""numpy restatement of ``DSP.wav_to_mel`` (utils/dsp.py:71-87,105-107)
(TEST INFRASTRUCTURE ONLY, see oracle/__init__).

The arithmetic of that method lives in the third-party dependency
librosa==0.7.2 (requirements.txt:2), which is neither vendored under
/root/reference nor installed here.  The functions below restate librosa's
published algorithm for the two calls the reference makes:

* ``librosa.stft(y, n_fft, hop_length, win_length)`` (utils/dsp.py:72-76) with
  its defaults window='hann', center=True, pad_mode='reflect', dtype=complex64:
  reflect-pad n_fft//2, frame, periodic Hann, float64 rFFT, store complex64.
* ``librosa.feature.melspectrogram(S=spec, sr, n_fft, n_mels, fmin, fmax)``
  (utils/dsp.py:78-84): ``S`` given -> ``power`` unused -> ``mel_basis @ S`` with
  ``librosa.filters.mel`` defaults htk=False, norm=1 (Slaney area norm), float32.

PARITY UNPINNED for librosa's internals: the only golden the reference holds
(tests/test_dsp.py:18-25 + tests/resources/test_mel.npy) needs an audio file
that ships inside librosa.  What *is* pinned here: frame count 1 + N//hop,
float32, clamp floor log(1e-5) (from that fixture) and agreement with
torchaudio's independent Slaney implementation (tests/test_oracle_dsp.py).
"""
from __future__ import annotations

import numpy as np


def hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1) -> (n_mels, 1+n_fft//2) float32."""
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_mag(y: np.ndarray, n_fft: int, hop: int, win: int) -> np.ndarray:
    """|librosa.stft| -> (1+n_fft//2, 1+N//hop) float32."""
    y = np.asarray(y)
    window = hann_periodic(win)
    if win < n_fft:  # librosa.util.pad_center
        lpad = (n_fft - win) // 2
        window = np.pad(window, (lpad, n_fft - win - lpad))
    yp = np.pad(y, n_fft // 2, mode='reflect')
    n_frames = 1 + (len(yp) - n_fft) // hop
    frames = np.lib.stride_tricks.as_strided(yp, shape=(n_fft, n_frames),
                                             strides=(yp.itemsize, hop * yp.itemsize))
    spec = np.fft.rfft(window[:, None] * frames, axis=0).astype(np.complex64)
    return np.abs(spec)


def wav_to_mel(y: np.ndarray, *, sample_rate=22050, n_fft=1024, hop_length=256, win_length=1024,
               num_mels=80, fmin=0, fmax=8000, normalize=True) -> np.ndarray:
    """utils/dsp.py:71-87 + normalize :105-107 -> (num_mels, 1+N//hop) float32."""
    spec = stft_mag(y, n_fft, hop_length, win_length)
    mel = np.dot(mel_filterbank(sample_rate, n_fft, num_mels, fmin, fmax), spec)
    if normalize:
        mel = np.log(np.clip(mel, a_min=1.e-5, a_max=None))
    return mel
