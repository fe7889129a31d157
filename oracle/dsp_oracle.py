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


# ----------------------------------------------------------------------------------------------------------------------
# DSP.griffinlim (utils/dsp.py:89-103) and DSP.trim_silence (utils/dsp.py:112-113): restated from librosa 0.7.2's
# published algorithms (the release is not on disk: PARITY UNPINNED against librosa itself, as for wav_to_mel).
# ----------------------------------------------------------------------------------------------------------------------
def stft(y: np.ndarray, n_fft: int = 1024, hop: int = 256, win: int = 1024) -> np.ndarray:
    """librosa.stft(y, n_fft, hop, win) (window='hann', center=True, pad_mode='reflect') -> complex64 (1+n_fft//2, F)."""
    window = hann_periodic(win)
    if win < n_fft:
        lpad = (n_fft - win) // 2
        window = np.pad(window, (lpad, n_fft - win - lpad))
    yp = np.pad(np.asarray(y), n_fft // 2, mode='reflect')
    n_frames = 1 + (len(yp) - n_fft) // hop
    frames = np.lib.stride_tricks.as_strided(yp, shape=(n_fft, n_frames), strides=(yp.itemsize, hop * yp.itemsize))
    return np.fft.rfft(window[:, None] * frames, axis=0).astype(np.complex64)


def window_sumsquare(n_frames: int, n_fft: int = 1024, hop: int = 256, win: int = 1024) -> np.ndarray:
    """librosa.filters.window_sumsquare('hann', n_frames, hop, win, n_fft, norm=None) -> float32 (n_fft + hop (F-1))."""
    win_sq = hann_periodic(win) ** 2
    if win < n_fft:
        lpad = (n_fft - win) // 2
        win_sq = np.pad(win_sq, (lpad, n_fft - win - lpad))
    x = np.zeros(n_fft + hop * (n_frames - 1), dtype=np.float32)
    for i in range(n_frames):
        x[i * hop:i * hop + n_fft] += win_sq.astype(np.float32)
    return x


def istft(S: np.ndarray, hop: int = 256, win: int = 1024) -> np.ndarray:
    """librosa.istft(S, hop, win) (window='hann', center=True, dtype=float32, length=None): windowed overlap-add of the
    inverse rFFTs, divided by the window sum-square where that exceeds tiny, centre padding cut."""
    n_fft = 2 * (S.shape[0] - 1)
    window = hann_periodic(win)
    if win < n_fft:
        lpad = (n_fft - win) // 2
        window = np.pad(window, (lpad, n_fft - win - lpad))
    n_frames = S.shape[1]
    y = np.zeros(n_fft + hop * (n_frames - 1), dtype=np.float32)
    for i in range(n_frames):
        y[i * hop:i * hop + n_fft] += (window * np.fft.irfft(S[:, i], n=n_fft)).astype(np.float32)
    wss = window_sumsquare(n_frames, n_fft, hop, win)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def griffinlim(S: np.ndarray, angles0: np.ndarray, n_iter: int = 32, hop: int = 256, win: int = 1024,
               momentum: float = 0.99) -> np.ndarray:
    """librosa.griffinlim(S, n_iter, hop, win) (momentum 0.99, init='random') with the random initial phases passed in:
    ``angles0`` = exp(2 pi i U) of the reference's rng.rand(*S.shape) (random_state=None upstream: not reproducible
    there, so the seeded phases are part of the test vector)."""
    n_fft = 2 * (S.shape[0] - 1)
    angles = angles0.astype(np.complex64).copy()
    rebuilt = 0.0
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft(S * angles, hop, win)
        rebuilt = stft(inverse, n_fft, hop, win)
        angles[:] = rebuilt - (momentum / (1 + momentum)) * tprev
        angles[:] /= np.abs(angles) + 1e-16
    return istft(S * angles, hop, win)


def mel_to_stft(mel_linear: np.ndarray, *, sample_rate=22050, n_fft=1024, fmin=0, fmax=8000) -> np.ndarray:
    """librosa.feature.inverse.mel_to_stft(M, power=1, sr, n_fft, fmin, fmax): non-negative least squares
    min_x 0.5 ||A x - M||^2, x >= 0, per block with L-BFGS-B started from the clipped least-squares solution
    (librosa.util.nnls -> _nnls_lbfgs_block).  The minimiser is not unique (80 x 513), so parity of an implementation is
    judged on the objective it reaches, not on x itself."""
    import scipy.optimize
    A = mel_filterbank(sample_rate, n_fft, mel_linear.shape[0], fmin, fmax).astype(mel_linear.dtype)
    B = mel_linear
    x_init = np.linalg.lstsq(A, B, rcond=None)[0]
    np.clip(x_init, 0, None, out=x_init)
    shape = x_init.shape

    def obj(x):
        x = x.reshape(shape)
        diff = np.dot(A, x) - B
        return 0.5 * np.sum(diff ** 2), np.dot(A.T, diff).flatten()

    x, _, _ = scipy.optimize.fmin_l_bfgs_b(obj, x_init, bounds=[(0, None)] * x_init.size, m=A.shape[1])
    return x.reshape(shape).astype(A.dtype)


def trim_silence(y: np.ndarray, top_db: float = 60, frame_length: int = 2048, hop_length: int = 512):
    """librosa.effects.trim(y, top_db, frame_length=2048, hop_length=512) -> (y[start:end], (start, end)):
    RMS per centred (reflect-padded) frame, power_to_db against the maximum, first / last frame above -top_db."""
    yp = np.pad(np.asarray(y, dtype=np.float32), frame_length // 2, mode='reflect')
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    frames = np.lib.stride_tricks.as_strided(yp, shape=(frame_length, n_frames),
                                             strides=(yp.itemsize, hop_length * yp.itemsize))
    mse = np.mean(np.abs(frames) ** 2, axis=0)
    amin = 1e-10
    db = 10.0 * np.log10(np.maximum(amin, mse)) - 10.0 * np.log10(np.maximum(amin, mse.max()))
    nonzero = np.flatnonzero(db > -top_db)
    if nonzero.size > 0:
        start = int(nonzero[0] * hop_length)
        end = min(len(y), int((nonzero[-1] + 1) * hop_length))
    else:
        start, end = 0, 0
    return y[start:end], (start, end)
