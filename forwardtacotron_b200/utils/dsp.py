"""Drop-in ``DSP`` for the STFT -> log-mel path (reference: utils/dsp.py).

Same constructor (all 19 ``dsp:`` keys of config.yaml:9-34 are kwargs, utils/dsp.py:14-57) and the
same ``wav_to_mel(y, normalize=True) -> (n_mels, 1 + len(y)//hop) float32`` (utils/dsp.py:71-87).
The arithmetic of the reference lives in librosa 0.7.2 on the CPU; here one fused sm_100a kernel does
framing, Hann window, 1024-point real FFT, magnitude, Slaney mel filterbank and log-clamp.

Beyond the reference surface, ``wav_to_mel_batch`` featurises many clips in ONE launch (the reference
reaches the same thing with a multiprocessing pool over files, preprocess.py:129-139) and
``wav_to_mel`` also accepts CUDA tensors so audio already resident in HBM is not copied back and forth.
The steps either side of that path are GPU kernels as well (csrc/griffin_lim.cu): ``griffinlim`` (mel -> NNLS linear
spectrogram -> 32 Griffin-Lim iterations, utils/dsp.py:89-103) and ``trim_silence`` (librosa.effects.trim,
utils/dsp.py:112-113).  File I/O, the webrtcvad long-silence trimmer and the mu-law helpers of the reference class are
CPU code outside this package.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Sequence, Union

import numpy as np
import torch

from .. import _lib


class ClipPlan:
    """Offsets of a packed batch of clips (``DSP.plan_clips``): samples and frames, host and device copies."""

    def __init__(self, offsets_host, frame_offsets_host, offsets, frame_offsets, total_frames: int):
        self.offsets_host, self.frame_offsets_host = offsets_host, frame_offsets_host
        self.offsets, self.frame_offsets, self.total_frames = offsets, frame_offsets, total_frames


class DSP:

    def __init__(self, num_mels: int, sample_rate: int, hop_length: int, win_length: int, n_fft: int, fmin: float,
                 fmax: float, peak_norm: bool, trim_start_end_silence: bool, trim_silence_top_db: int,
                 pitch_max_freq: int, trim_long_silences: bool, vad_sample_rate: int, vad_window_length: float,
                 vad_moving_average_width: float, vad_max_silence_length: int, bits: int, mu_law: bool,
                 voc_mode: str, device: Union[str, torch.device, None] = None) -> None:
        self.n_mels = num_mels
        self.sample_rate = sample_rate
        self.hop_length = hop_length
        self.win_length = win_length
        self.n_fft = n_fft
        self.fmin = fmin
        self.fmax = fmax
        self.should_peak_norm = peak_norm
        self.should_trim_start_end_silence = trim_start_end_silence
        self.should_trim_long_silences = trim_long_silences
        self.trim_silence_top_db = trim_silence_top_db
        self.pitch_max_freq = pitch_max_freq
        self.vad_sample_rate = vad_sample_rate
        self.vad_window_length = vad_window_length
        self.vad_moving_average_width = vad_moving_average_width
        self.vad_max_silence_length = vad_max_silence_length
        self.bits = bits
        self.mu_law = mu_law
        self.voc_mode = voc_mode
        self._device = torch.device(device) if device is not None else None
        self._handles: Dict[int, C.c_void_p] = {}

    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> 'DSP':
        return DSP(**config['dsp'])

    # ------------------------------------------------------------------ native plumbing
    def _handle(self, device: torch.device) -> C.c_void_p:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            cfg = _lib.MelConfig(int(self.sample_rate), int(self.n_fft), int(self.hop_length), int(self.win_length),
                                 int(self.n_mels), float(self.fmin), float(self.fmax))
            h = C.c_void_p()
            _lib.check(_lib.lib().ftb_mel_create(C.byref(cfg), idx, C.byref(h)))
            self._handles[idx] = h
        return h

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().ftb_mel_destroy(h)
        except Exception:
            pass

    def _default_device(self) -> torch.device:
        if self._device is not None:
            return self._device
        if not torch.cuda.is_available():
            raise RuntimeError('DSP.wav_to_mel runs on sm_100a GPUs only and no CUDA device is available '
                               '(there is no CPU fallback)')
        return torch.device('cuda', torch.cuda.current_device())

    def mel_filterbank(self) -> np.ndarray:
        """The (n_mels, 1 + n_fft//2) float32 Slaney filterbank the kernel uses."""
        h = self._handle(self._default_device())
        fb = np.empty((self.n_mels, 1 + self.n_fft // 2), dtype=np.float32)
        _lib.check(_lib.lib().ftb_mel_filterbank(h, fb.ctypes.data_as(C.c_void_p)))
        return fb

    # ------------------------------------------------------------------ the hot path
    def plan_clips(self, clip_offsets: torch.Tensor, device) -> 'ClipPlan':
        """Frame bookkeeping of a packed batch of clips, done ONCE per clip layout: sample / frame offsets on the host
        and on the device.  A plan can be reused for every ``wav_to_mel_packed`` call over audio with the same clip
        boundaries (e.g. the passes of a benchmark, or Griffin-Lim iterations), which then launch the kernel with no
        host-side arithmetic and no H2D copy at all."""
        offs_cpu = clip_offsets.detach().to('cpu', torch.int64)
        lens = offs_cpu[1:] - offs_cpu[:-1]
        if len(lens) == 0 or int(lens.min()) < 1:
            raise ValueError('empty clip')
        frames = 1 + lens // self.hop_length  # librosa.stft with center=True
        fo_cpu = torch.zeros(len(lens) + 1, dtype=torch.int64)
        fo_cpu[1:] = torch.cumsum(frames, 0)
        both = torch.cat([offs_cpu, fo_cpu]).pin_memory().to(device, non_blocking=True)   # one H2D copy
        n = len(offs_cpu)
        return ClipPlan(offs_cpu, fo_cpu, both[:n], both[n:], int(fo_cpu[-1]))

    def wav_to_mel_packed(self, audio: torch.Tensor, clip_offsets, normalize: bool = True, out=None):
        """audio: flat float32 CUDA tensor holding all clips back to back; clip_offsets: (n+1) int64 tensor (CPU or
        CUDA) or a ``ClipPlan`` from ``plan_clips``.  Returns (flat output, frame_offsets on the host): clip i is
        ``out[80*fo[i]:80*fo[i+1]].view(80, -1)``."""
        if not audio.is_cuda:
            raise RuntimeError('wav_to_mel_packed expects audio resident on the GPU')
        dev = audio.device
        audio = audio.to(torch.float32).contiguous()
        plan = clip_offsets if isinstance(clip_offsets, ClipPlan) else self.plan_clips(clip_offsets, dev)
        if int(plan.offsets_host[-1]) > audio.numel():
            raise ValueError('clip offsets reach beyond the audio buffer')
        if out is None:
            out = torch.empty(self.n_mels * plan.total_frames, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ftb_mel_run(self._handle(dev), _lib.ptr(audio), _lib.ptr(plan.offsets), _lib.ptr(plan.frame_offsets),
                                              len(plan.offsets_host) - 1, plan.total_frames, _lib.ptr(out),
                                              int(bool(normalize)), _lib.current_stream(dev)))
        return out, plan.frame_offsets_host

    def wav_to_mel_batch(self, clips: Sequence[Union[np.ndarray, torch.Tensor]], normalize: bool = True) -> List:
        """Many clips, one launch.  Returns a list of (n_mels, frames_i) arrays / tensors (numpy in -> numpy out)."""
        dev = self._default_device()
        as_numpy = not isinstance(clips[0], torch.Tensor)
        ts = [torch.as_tensor(np.ascontiguousarray(c), dtype=torch.float32) if not isinstance(c, torch.Tensor)
              else c.to(torch.float32) for c in clips]
        dev = ts[0].device if ts[0].is_cuda else dev
        offs = torch.zeros(len(ts) + 1, dtype=torch.int64)
        offs[1:] = torch.cumsum(torch.tensor([t.numel() for t in ts]), 0)
        flat = torch.cat([t.reshape(-1) for t in ts])
        if not flat.is_cuda:
            flat = flat.pin_memory().to(dev, non_blocking=True)
        out, fo = self.wav_to_mel_packed(flat, offs, normalize)
        res = []
        host = out.cpu().numpy() if as_numpy else None
        for i in range(len(ts)):
            a, b = self.n_mels * int(fo[i]), self.n_mels * int(fo[i + 1])
            res.append(host[a:b].reshape(self.n_mels, -1) if as_numpy else out[a:b].view(self.n_mels, -1))
        return res

    def wav_to_mel(self, y: Union[np.ndarray, torch.Tensor], normalize=True) -> Union[np.ndarray, torch.Tensor]:
        return self.wav_to_mel_batch([y], normalize)[0]

    # ------------------------------------------------------------------ the inverse path
    def mel_to_stft(self, mel, denormalize: bool = False, iters: int = 64):
        """``librosa.feature.inverse.mel_to_stft(M, power=1, ...)``: (n_mels, F) mel magnitudes (log-mels when
        ``denormalize``) -> (1 + n_fft//2, F) non-negative linear magnitudes.  numpy in -> numpy out."""
        dev = self._default_device()
        as_numpy = not isinstance(mel, torch.Tensor)
        m = torch.as_tensor(np.ascontiguousarray(mel), dtype=torch.float32) if as_numpy else mel.to(torch.float32)
        m = m.to(dev).contiguous()
        if m.dim() != 2 or m.shape[0] != self.n_mels:
            raise ValueError(f'mel must be ({self.n_mels}, frames)')
        F = int(m.shape[1])
        S = torch.empty((1 + self.n_fft // 2, F), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ftb_mel_to_stft(self._handle(dev), _lib.ptr(m), F, int(bool(denormalize)), int(iters),
                                                  _lib.ptr(S), _lib.current_stream(dev)))
        return S.cpu().numpy() if as_numpy else S

    def griffinlim_from_stft(self, S, phase_u, n_iter: int = 32, momentum: float = 0.99):
        """``librosa.griffinlim(S, n_iter, hop_length, win_length)`` from the initial phases ``exp(2 pi i phase_u)``
        (``phase_u``: uniform [0, 1) of S's shape -- upstream draws them from an unseeded RNG)."""
        dev = self._default_device()
        as_numpy = not isinstance(S, torch.Tensor)
        St = torch.as_tensor(np.ascontiguousarray(S), dtype=torch.float32).to(dev).contiguous() if as_numpy \
            else S.to(dev, torch.float32).contiguous()
        u = torch.as_tensor(np.ascontiguousarray(phase_u), dtype=torch.float32).to(dev).contiguous() \
            if not isinstance(phase_u, torch.Tensor) else phase_u.to(dev, torch.float32).contiguous()
        if St.shape != u.shape or St.dim() != 2 or St.shape[0] != 1 + self.n_fft // 2:
            raise ValueError('S and phase_u must both be (1 + n_fft//2, frames)')
        F = int(St.shape[1])
        if F < 2:
            raise ValueError('Griffin-Lim needs at least 2 frames')
        lib = _lib.lib()
        ws = torch.empty(int(lib.ftb_griffinlim_workspace_bytes(F)), dtype=torch.uint8, device=dev)
        wav = torch.empty(self.hop_length * (F - 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ftb_griffinlim(self._handle(dev), _lib.ptr(St), _lib.ptr(u), F, int(n_iter), float(momentum),
                                          _lib.ptr(wav), _lib.ptr(ws), ws.numel(), _lib.current_stream(dev)))
        return wav.cpu().numpy() if as_numpy else wav

    def griffinlim(self, mel, n_iter: int = 32, seed=None):
        """utils/dsp.py:89-103: log-mel (n_mels, F) -> waveform of hop * (F - 1) samples.  ``seed`` fixes the random
        initial phases (None = fresh ones per call, like upstream)."""
        dev = self._default_device()
        S = self.mel_to_stft(mel, denormalize=True)
        shape = S.shape
        g = None
        if seed is not None:
            g = torch.Generator(device=dev)
            g.manual_seed(int(seed))
        u = torch.rand(shape, generator=g, device=dev, dtype=torch.float32)
        return self.griffinlim_from_stft(S, u if isinstance(S, torch.Tensor) else u.cpu().numpy(), n_iter)

    def trim_bounds(self, audio: torch.Tensor, clip_offsets: torch.Tensor) -> torch.Tensor:
        """librosa.effects.trim over clips packed back to back on the GPU -> (n_clips, 2) int64 [start, end) per clip."""
        if not audio.is_cuda:
            raise RuntimeError('trim_bounds expects audio resident on the GPU')
        dev = audio.device
        offs_cpu = clip_offsets.detach().to('cpu', torch.int64)
        lens = offs_cpu[1:] - offs_cpu[:-1]
        if len(lens) == 0 or int(lens.min()) < 1:
            raise ValueError('empty clip')
        n, mx, hop = len(lens), int(lens.max()), 512
        offs = offs_cpu.pin_memory().to(dev, non_blocking=True)
        ws = torch.empty(n * (1 + mx // hop), dtype=torch.float32, device=dev)
        bounds = torch.empty((n, 2), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ftb_trim_silence(_lib.ptr(audio.to(torch.float32).contiguous()), _lib.ptr(offs), n, mx,
                                                   float(self.trim_silence_top_db), 2048, hop, _lib.ptr(bounds),
                                                   _lib.ptr(ws), ws.numel() * 4, _lib.current_stream(dev)))
        return bounds

    def trim_silence(self, wav):
        """utils/dsp.py:112-113: cut leading / trailing frames more than ``trim_silence_top_db`` dB below the loudest."""
        dev = self._default_device()
        as_numpy = not isinstance(wav, torch.Tensor)
        t = torch.as_tensor(np.ascontiguousarray(wav), dtype=torch.float32) if as_numpy else wav.to(torch.float32)
        d = t.to(dev).contiguous()
        b = self.trim_bounds(d, torch.tensor([0, d.numel()])).cpu()
        return wav[int(b[0, 0]):int(b[0, 1])]

    def normalize(self, mel: np.ndarray) -> np.ndarray:
        mel = np.clip(mel, a_min=1.e-5, a_max=None)
        return np.log(mel)

    def denormalize(self, mel: np.ndarray) -> np.ndarray:
        return np.exp(mel)
