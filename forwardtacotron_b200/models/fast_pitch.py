"""Drop-in ``FastPitch`` for inference on B200 (reference: models/fast_pitch.py).

Transformer (post-LN FFT blocks) instead of CBHG/RNN.  Parameter tree and names follow the
reference (277 state_dict entries incl. the ``pos_encoder.pe`` buffers, models/fast_pitch.py:16-236);
``generate`` has the reference semantics (:286-340): predictors and postnet run WITHOUT a padding
mask, the prenet masks keys where ``x == 0``, 'mel_post' is the same tensor as 'mel', and a
sequence longer than the 5000-row positional table raises.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Any, Callable, Dict, Union

import torch
import torch.nn as nn

from .. import _lib
from ..utils.text import NUM_PHONEMES
from .common_layers import LengthRegulator, NativeModel, _require_cuda


class PositionalEncoding(nn.Module):
    def __init__(self, d_model: int, dropout=0.1, max_len=5000) -> None:
        super().__init__()
        self.scale = nn.Parameter(torch.ones(1))
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer('pe', pe.unsqueeze(1))  # (max_len, 1, d_model) as in the reference


class FFTBlock(nn.Module):
    def __init__(self, d_model: int, nhead: int, conv1_kernel: int, conv2_kernel: int, d_fft: int,
                 dropout: float = 0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.conv1 = nn.Conv1d(d_model, d_fft, conv1_kernel, stride=1, padding=conv1_kernel // 2)
        self.conv2 = nn.Conv1d(d_fft, d_model, conv2_kernel, stride=1, padding=conv2_kernel // 2)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class ForwardTransformer(nn.Module):
    def __init__(self, d_model: int, d_fft: int, layers: int, heads: int, conv1_kernel: int, conv2_kernel: int,
                 dropout: float = 0.1) -> None:
        super().__init__()
        self.d_model = d_model
        self.pos_encoder = PositionalEncoding(d_model, dropout)
        # the reference deep-copies ONE initialised block (models/fast_pitch.py:115); keep that init behaviour
        proto = FFTBlock(d_model, heads, conv1_kernel, conv2_kernel, d_fft, dropout)
        self.layers = nn.ModuleList()
        for _ in range(layers):
            blk = FFTBlock(d_model, heads, conv1_kernel, conv2_kernel, d_fft, dropout)
            blk.load_state_dict(proto.state_dict())
            self.layers.append(blk)
        self.norm = nn.LayerNorm(d_model)


class SeriesPredictor(nn.Module):
    def __init__(self, num_chars: int, d_model: int, n_heads: int, d_fft: int, layers: int, conv1_kernel: int,
                 conv2_kernel: int, dropout=0.1):
        super().__init__()
        self.embedding = nn.Embedding(num_chars, d_model)
        self.transformer = ForwardTransformer(heads=n_heads, dropout=dropout, d_model=d_model, d_fft=d_fft,
                                              conv1_kernel=conv1_kernel, conv2_kernel=conv2_kernel, layers=layers)
        self.lin = nn.Linear(d_model, 1)


class FastPitch(NativeModel):
    _create_fn = 'ftb_fp_create'
    _destroy_fn = 'ftb_fp_destroy'

    def __init__(self, num_chars: int, durpred_dropout: float, durpred_d_model: int, durpred_n_heads: int,
                 durpred_layers: int, durpred_d_fft: int, pitch_dropout: float, pitch_d_model: int,
                 pitch_n_heads: int, pitch_layers: int, pitch_d_fft: int, energy_dropout: float,
                 energy_d_model: int, energy_n_heads: int, energy_layers: int, energy_d_fft: int,
                 pitch_strength: float, energy_strength: float, d_model: int, conv1_kernel: int, conv2_kernel: int,
                 prenet_layers: int, prenet_heads: int, prenet_fft: int, prenet_dropout: float, postnet_layers: int,
                 postnet_heads: int, postnet_fft: int, postnet_dropout: float, n_mels: int, padding_value=-11.5129):
        super().__init__()
        self.padding_value = padding_value
        self.pitch_strength = pitch_strength
        self.energy_strength = energy_strength
        loc = dict(locals())
        self._dims = {k: int(loc[k]) for k in _lib.FP_INT_FIELDS}
        self.lr = LengthRegulator()
        self.dur_pred = SeriesPredictor(num_chars, durpred_d_model, durpred_n_heads, durpred_d_fft, durpred_layers,
                                        conv1_kernel, conv2_kernel, durpred_dropout)
        self.pitch_pred = SeriesPredictor(num_chars, pitch_d_model, pitch_n_heads, pitch_d_fft, pitch_layers,
                                          conv1_kernel, conv2_kernel, pitch_dropout)
        self.energy_pred = SeriesPredictor(num_chars, energy_d_model, energy_n_heads, energy_d_fft, energy_layers,
                                           conv1_kernel, conv2_kernel, energy_dropout)
        self.embedding = nn.Embedding(num_embeddings=num_chars, embedding_dim=d_model)
        self.prenet = ForwardTransformer(d_model=d_model, d_fft=prenet_fft, layers=prenet_layers, heads=prenet_heads,
                                         conv1_kernel=conv1_kernel, conv2_kernel=conv2_kernel, dropout=prenet_dropout)
        self.postnet = ForwardTransformer(d_model=d_model, d_fft=postnet_fft, layers=postnet_layers,
                                          heads=postnet_heads, conv1_kernel=conv1_kernel, conv2_kernel=conv2_kernel,
                                          dropout=postnet_dropout)
        self.lin = nn.Linear(d_model, n_mels)
        self.register_buffer('step', torch.zeros(1, dtype=torch.long))
        self.pitch_proj = nn.Conv1d(1, d_model, kernel_size=3, padding=1)
        self.energy_proj = nn.Conv1d(1, d_model, kernel_size=3, padding=1)

    def __repr__(self):
        return f'FastPitch, num params: {sum(p.numel() for p in self.parameters())}'

    def get_step(self) -> int:
        return self.step.data.item()

    def _config_struct(self) -> _lib.FpConfig:
        cfg = _lib.FpConfig()
        for k in _lib.FP_INT_FIELDS:
            setattr(cfg, k, self._dims[k])
        cfg.pitch_strength = float(self.pitch_strength)
        cfg.energy_strength = float(self.energy_strength)
        cfg.gemm_mode = int(self.gemm_mode)
        return cfg

    def _check_tokens(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, 'generate')
        if x.dim() != 2 or x.dtype != torch.long:
            raise TypeError('x must be an int64 tensor of shape (B, T)')
        return x.contiguous()

    def _max_len(self) -> int:
        return int(self.prenet.pos_encoder.pe.shape[0])

    def _workspace_for(self, handle, B, T, L, device):
        n = _lib.lib().ftb_fp_workspace_bytes(handle, B, T, L)
        if n < 0:
            raise _lib.FtbError(int(n), 'ftb_fp_workspace_bytes failed')
        return self._get_workspace(n, device)

    def _check_len(self, S: int) -> None:
        if S > self._max_len():  # the reference fails in the pe broadcast (models/fast_pitch.py:33)
            raise RuntimeError(f'The size of tensor a ({S}) must match the size of tensor b ({self._max_len()}) '
                               'at non-singleton dimension 0')

    def predict(self, x: torch.Tensor, alpha: float = 1.0):
        x = self._check_tokens(x)
        lib, dev = _lib.lib(), x.device
        B, T = x.shape
        self._check_len(T)
        h = self._get_handle(dev)
        ws = self._workspace_for(h, B, T, 0, dev)
        dur = torch.empty((B, T), dtype=torch.float32, device=dev)
        pitch = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
        energy = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ftb_fp_predict(h, _lib.ptr(x), B, T, float(alpha), _lib.ptr(dur), _lib.ptr(pitch),
                                          _lib.ptr(energy), _lib.ptr(ws), ws.numel(), _lib.current_stream(dev)))
        return dur, pitch, energy

    def synthesize(self, x, dur_hat, pitch_hat, energy_hat, mel_post_alloc=None) -> Dict[str, torch.Tensor]:
        x = self._check_tokens(x)
        lib, dev = _lib.lib(), x.device
        B, T = x.shape
        h = self._get_handle(dev)
        if dur_hat.dtype != torch.float32 or not dur_hat.is_contiguous() or dur_hat.shape != (B, T):
            raise TypeError('dur_hat must be a contiguous float32 (B, T) tensor')
        pitch_c = pitch_hat.to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
        energy_c = energy_hat.to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
        cum, total = LengthRegulator.plan(dur_hat)
        L = int(total.max().item())
        if L <= 0:
            raise RuntimeError('all rounded durations are zero: nothing to synthesize')
        self._check_len(L)
        ws = self._workspace_for(h, B, T, L, dev)
        # optional caller-supplied output memory (may be peer-mapped, utils/peer_window.py), as in ForwardTacotron
        mel = torch.empty((B, self._dims['n_mels'], L), dtype=torch.float32, device=dev) if mel_post_alloc is None \
            else mel_post_alloc(B, self._dims['n_mels'], L)
        if not (mel.is_cuda and mel.dtype == torch.float32 and mel.is_contiguous()
                and tuple(mel.shape) == (B, self._dims['n_mels'], L)):
            raise TypeError('mel_post_alloc must return a contiguous float32 CUDA tensor of shape (B, n_mels, L)')
        with torch.cuda.device(dev):
            _lib.check(lib.ftb_fp_synthesize(h, _lib.ptr(x), _lib.ptr(cum), _lib.ptr(pitch_c), _lib.ptr(energy_c),
                                             B, T, L, _lib.ptr(mel), _lib.ptr(ws), ws.numel(),
                                             _lib.current_stream(dev)))
        return {'mel': mel, 'mel_post': mel, 'dur': dur_hat, 'pitch': pitch_hat, 'energy': energy_hat,
                'mel_len': total}

    def generate(self, x: torch.Tensor, alpha=1.0,
                 pitch_function: Callable[[torch.Tensor], torch.Tensor] = lambda x: x,
                 energy_function: Callable[[torch.Tensor], torch.Tensor] = lambda x: x,
                 mel_post_alloc=None) -> Dict[str, torch.Tensor]:
        self.eval()
        with torch.no_grad():
            dur_hat, pitch_hat, energy_hat = self.predict(x, alpha)
            pitch_hat = pitch_function(pitch_hat)
            energy_hat = energy_function(energy_hat)
            return self.synthesize(x, dur_hat, pitch_hat, energy_hat, mel_post_alloc)

    def forward(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Teacher-forced forward in EVAL mode (models/fast_pitch.py:243-283): predictors and prenet with the token
        padding mask, durations / pitch / energy of the batch drive the synthesis, the postnet masks the frames past
        ``mel_len`` as keys, outputs cut / padded to ``mel.size(2)`` with ``padding_value``; 'mel_post' is 'mel'.
        Training mode is outside the path this package implements."""
        if self.training:
            raise NotImplementedError('training-mode forward() is outside the hot path this package implements '
                                      '(SURVEY 8f-4); call .eval() for the teacher-forced inference pass')
        with torch.no_grad():
            x = self._check_tokens(batch['x'])
            lib, dev = _lib.lib(), x.device
            B, T = x.shape
            self._check_len(T)
            dur = batch['dur']
            if dur.dtype != torch.float32 or not dur.is_contiguous() or dur.device != dev:
                dur = dur.to(device=dev, dtype=torch.float32).contiguous()
            pitch = batch['pitch'].to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
            energy = batch['energy'].to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
            cum, total = LengthRegulator.plan(dur)
            L = int(total.max().item())
            if L <= 0:
                raise RuntimeError('all rounded durations are zero: nothing to synthesize')
            self._check_len(L)
            mel_lens = batch['mel_len'].to(dev)
            frame_mask = (torch.arange(L, device=dev)[None, :] < mel_lens[:, None]).to(torch.int64).contiguous()
            h = self._get_handle(dev)
            ws = self._workspace_for(h, B, T, L, dev)
            dur_hat = torch.empty((B, T), dtype=torch.float32, device=dev)
            pitch_hat = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
            energy_hat = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
            mel = torch.empty((B, self._dims['n_mels'], L), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.ftb_fp_forward_eval(h, _lib.ptr(x), _lib.ptr(cum), _lib.ptr(pitch), _lib.ptr(energy),
                                                   _lib.ptr(frame_mask), B, T, L, _lib.ptr(dur_hat), _lib.ptr(pitch_hat),
                                                   _lib.ptr(energy_hat), _lib.ptr(mel), _lib.ptr(ws), ws.numel(),
                                                   _lib.current_stream(dev)))
            max_len = int(batch['mel'].size(2))
            mel = mel[:, :, :max_len]
            mel = torch.nn.functional.pad(mel, [0, max_len - mel.size(2), 0, 0], 'constant', self.padding_value)
            return {'mel': mel, 'mel_post': mel, 'dur': dur_hat, 'pitch': pitch_hat, 'energy': energy_hat}

    def last_launch_count(self) -> int:
        return int(_lib.lib().ftb_fp_last_launch_count(self._handle)) if self._handle is not None else 0

    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> 'FastPitch':
        model_config = config['fast_pitch']['model']
        model_config['num_chars'] = NUM_PHONEMES
        model_config['n_mels'] = config['dsp']['num_mels']
        return FastPitch(**model_config)

    @classmethod
    def from_checkpoint(cls, path: Union[Path, str]) -> 'FastPitch':
        checkpoint = torch.load(path, map_location=torch.device('cpu'))
        model = FastPitch.from_config(checkpoint['config'])
        model.load_state_dict(checkpoint['model'])
        return model
