"""Drop-in ``ForwardTacotron`` for inference on B200 (reference: models/forward_tacotron.py).

Same constructor kwargs (config.yaml:78-106), same parameter / buffer names (so the
322-entry ``state_dict`` of an upstream checkpoint loads with ``strict=True``), same
``from_config`` / ``from_checkpoint`` / ``generate`` / ``get_step`` surface
(models/forward_tacotron.py:244-268,286-287,338-350).  ``generate`` runs entirely in the
sm_100a extension:

    stage A  ftb_ft_predict      dur / pitch / energy predictors, duration fallback (:251-262)
    host     pitch_function / energy_function callbacks on real torch tensors (:259,:263)
    plan     ftb_length_plan + ONE device->host read of the B frame counts (output size)
    stage B  ftb_ft_synthesize   prenet, conditioning, LengthRegulator, LSTM, postnet (:289-330)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Any, Callable, Dict, Union

import torch
import torch.nn as nn

from .. import _lib
from ..utils.text import NUM_PHONEMES
from .common_layers import CBHG, BatchNormConv, LengthRegulator, NativeModel, _require_cuda


class SeriesPredictor(nn.Module):
    """embed -> 3 x (conv5, ReLU, BN) -> biGRU -> Linear(2H -> 1)   (models/forward_tacotron.py:14-55)"""

    def __init__(self, num_chars, emb_dim=64, conv_dims=256, rnn_dims=64, dropout=0.5):
        super().__init__()
        self.embedding = nn.Embedding(num_chars, emb_dim)
        self.convs = nn.ModuleList([BatchNormConv(emb_dim if i == 0 else conv_dims, conv_dims, 5, relu=True)
                                    for i in range(3)])
        self.rnn = nn.GRU(conv_dims, rnn_dims, batch_first=True, bidirectional=True)
        self.lin = nn.Linear(2 * rnn_dims, 1)
        self.dropout = dropout


class ForwardTacotron(NativeModel):
    _create_fn = 'ftb_ft_create'
    _destroy_fn = 'ftb_ft_destroy'
    _PREDICTORS = {'dur_pred': 0, 'pitch_pred': 1, 'energy_pred': 2}

    def __init__(self, embed_dims: int, series_embed_dims: int, num_chars: int, durpred_conv_dims: int,
                 durpred_rnn_dims: int, durpred_dropout: float, pitch_conv_dims: int, pitch_rnn_dims: int,
                 pitch_dropout: float, pitch_strength: float, energy_conv_dims: int, energy_rnn_dims: int,
                 energy_dropout: float, energy_strength: float, rnn_dims: int, prenet_dims: int, prenet_k: int,
                 postnet_num_highways: int, prenet_dropout: float, postnet_dims: int, postnet_k: int,
                 prenet_num_highways: int, postnet_dropout: float, n_mels: int, padding_value=-11.5129):
        super().__init__()
        self.rnn_dims = rnn_dims
        self.padding_value = padding_value
        self.pitch_strength = pitch_strength
        self.energy_strength = energy_strength
        self._dims = dict(num_chars=num_chars, embed_dims=embed_dims, series_embed_dims=series_embed_dims,
                          durpred_conv_dims=durpred_conv_dims, durpred_rnn_dims=durpred_rnn_dims,
                          pitch_conv_dims=pitch_conv_dims, pitch_rnn_dims=pitch_rnn_dims,
                          energy_conv_dims=energy_conv_dims, energy_rnn_dims=energy_rnn_dims, rnn_dims=rnn_dims,
                          prenet_dims=prenet_dims, prenet_k=prenet_k, prenet_num_highways=prenet_num_highways,
                          postnet_dims=postnet_dims, postnet_k=postnet_k,
                          postnet_num_highways=postnet_num_highways, n_mels=n_mels)
        # registration order follows the reference so state_dict() iterates identically
        self.register_buffer('step', torch.zeros(1, dtype=torch.long))
        self.embedding = nn.Embedding(num_chars, embed_dims)
        self.prenet = CBHG(K=prenet_k, in_channels=embed_dims, channels=prenet_dims,
                           proj_channels=[prenet_dims, embed_dims], num_highways=prenet_num_highways,
                           dropout=prenet_dropout)
        self.pitch_pred = SeriesPredictor(num_chars, series_embed_dims, pitch_conv_dims, pitch_rnn_dims, pitch_dropout)
        self.energy_pred = SeriesPredictor(num_chars, series_embed_dims, energy_conv_dims, energy_rnn_dims,
                                           energy_dropout)
        self.pitch_proj = nn.Conv1d(1, 2 * prenet_dims, kernel_size=3, padding=1)
        self.energy_proj = nn.Conv1d(1, 2 * prenet_dims, kernel_size=3, padding=1)
        self.lr = LengthRegulator()
        self.dur_pred = SeriesPredictor(num_chars, series_embed_dims, durpred_conv_dims, durpred_rnn_dims,
                                        durpred_dropout)
        self.lstm = nn.LSTM(2 * prenet_dims, rnn_dims, batch_first=True, bidirectional=True)
        self.lin = nn.Linear(2 * rnn_dims, n_mels)
        self.postnet = CBHG(K=postnet_k, in_channels=n_mels, channels=postnet_dims,
                            proj_channels=[postnet_dims, n_mels], num_highways=postnet_num_highways,
                            dropout=postnet_dropout)
        self.post_proj = nn.Linear(2 * postnet_dims, n_mels, bias=False)

    def __repr__(self):
        return f'ForwardTacotron, num params: {sum(p.numel() for p in self.parameters())}'

    def get_step(self) -> int:
        return self.step.data.item()

    # ------------------------------------------------------------------ native plumbing
    def _config_struct(self) -> _lib.FtConfig:
        cfg = _lib.FtConfig()
        for k in _lib.FT_INT_FIELDS:
            setattr(cfg, k, int(self._dims[k]))
        cfg.pitch_strength = float(self.pitch_strength)
        cfg.energy_strength = float(self.energy_strength)
        cfg.gemm_mode = int(self.gemm_mode)
        return cfg

    def _check_tokens(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, 'generate')
        if x.dim() != 2 or x.dtype != torch.long:
            raise TypeError('x must be an int64 tensor of shape (B, T)')
        return x.contiguous()

    def _workspace_for(self, handle, B: int, T: int, L: int, device) -> torch.Tensor:
        n = _lib.lib().ftb_ft_workspace_bytes(handle, B, T, L)
        if n < 0:
            raise _lib.FtbError(int(n), 'ftb_ft_workspace_bytes failed')
        return self._get_workspace(n, device)

    # ------------------------------------------------------------------ stages
    def _check_lengths(self, x: torch.Tensor, lengths) -> torch.Tensor:
        lens = torch.as_tensor(lengths).to(device=x.device, dtype=torch.int32).contiguous()
        if lens.shape != (x.shape[0],):
            raise TypeError('lengths must have one entry per row of x')
        return lens

    def predict(self, x: torch.Tensor, alpha: float = 1.0, lengths=None):
        """Stage A -> (dur (B,T), pitch (B,1,T), energy (B,1,T)), fallback applied.  ``lengths`` (B,) token counts:
        ragged batch, every row computed as if alone, outputs zero beyond its length (see ``generate_ragged``)."""
        x = self._check_tokens(x)
        lib, dev = _lib.lib(), x.device
        B, T = x.shape
        h = self._get_handle(dev)
        ws = self._workspace_for(h, B, T, 0, dev)
        dur = torch.empty((B, T), dtype=torch.float32, device=dev)
        pitch = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
        energy = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if lengths is None:
                _lib.check(lib.ftb_ft_predict(h, _lib.ptr(x), B, T, float(alpha), _lib.ptr(dur), _lib.ptr(pitch),
                                              _lib.ptr(energy), _lib.ptr(ws), ws.numel(), _lib.current_stream(dev)))
            else:
                lens = self._check_lengths(x, lengths)
                _lib.check(lib.ftb_ft_predict_ragged(h, _lib.ptr(x), _lib.ptr(lens), B, T, float(alpha), _lib.ptr(dur),
                                                     _lib.ptr(pitch), _lib.ptr(energy), _lib.ptr(ws), ws.numel(),
                                                     _lib.current_stream(dev)))
        return dur, pitch, energy

    def synthesize(self, x: torch.Tensor, dur_hat: torch.Tensor, pitch_hat: torch.Tensor,
                   energy_hat: torch.Tensor, mel_post_alloc=None, lengths=None) -> Dict[str, torch.Tensor]:
        """Stage B == the reference's ``_generate_mel`` (:289-330); clamps ``dur_hat`` in place.
        ``lengths``: ragged batch (see ``generate_ragged``); frames of row b beyond ``mel_len[b]`` are padding.

        ``mel_post_alloc(B, n_mels, L) -> float32 tensor`` (or utils/peer_window.PeerSlot; optional) supplies the memory ``mel_post`` is written to
        by the last GEMM's epilogue.  It may live on ANOTHER GPU of the node (peer-mapped, utils/peer_window.py): the
        final gather of a sharded run then costs no extra pass."""
        x = self._check_tokens(x)
        lib, dev = _lib.lib(), x.device
        B, T = x.shape
        h = self._get_handle(dev)
        if dur_hat.dtype != torch.float32 or not dur_hat.is_contiguous() or dur_hat.shape != (B, T):
            raise TypeError('dur_hat must be a contiguous float32 (B, T) tensor')
        pitch_c = pitch_hat.to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
        energy_c = energy_hat.to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
        lens = None
        if lengths is not None:
            # the callbacks may have written to the padded positions (e.g. e + 0.1): the conditioning convs of a solo
            # run see zeros there, and padded tokens must not expand into frames
            lens = self._check_lengths(x, lengths)
            if pitch_c.data_ptr() == pitch_hat.data_ptr():
                pitch_c = pitch_c.clone()
            if energy_c.data_ptr() == energy_hat.data_ptr():
                energy_c = energy_c.clone()
            with torch.cuda.device(dev):
                for t in (pitch_c, energy_c, dur_hat):
                    _lib.check(lib.ftb_zero_tail_rows(_lib.ptr(t), B, T, 4, _lib.ptr(lens), _lib.current_stream(dev)))
        cum, total = LengthRegulator.plan(dur_hat)
        L = int(total.max().item())  # D2H: sizes the outputs
        if L <= 0:
            raise RuntimeError('all rounded durations are zero: nothing to synthesize')
        ws = self._workspace_for(h, B, T, L, dev)
        n_mels = self._dims['n_mels']
        mel = torch.empty((B, n_mels, L), dtype=torch.float32, device=dev)
        if mel_post_alloc is None:
            mel_post = torch.empty((B, n_mels, L), dtype=torch.float32, device=dev)
        else:
            mel_post = mel_post_alloc(B, n_mels, L)
            if not (mel_post.is_cuda and mel_post.dtype == torch.float32 and mel_post.is_contiguous()
                    and tuple(mel_post.shape) == (B, n_mels, L)):
                raise TypeError('mel_post_alloc must return a contiguous float32 CUDA tensor of shape (B, n_mels, L)')
        with torch.cuda.device(dev):
            if lens is None:
                _lib.check(lib.ftb_ft_synthesize(h, _lib.ptr(x), _lib.ptr(cum), _lib.ptr(pitch_c), _lib.ptr(energy_c),
                                                 B, T, L, _lib.ptr(mel), _lib.ptr(mel_post), _lib.ptr(ws), ws.numel(),
                                                 _lib.current_stream(dev)))
            else:
                _lib.check(lib.ftb_ft_synthesize_ragged(h, _lib.ptr(x), _lib.ptr(lens), _lib.ptr(cum), _lib.ptr(pitch_c),
                                                        _lib.ptr(energy_c), _lib.ptr(total), B, T, L, _lib.ptr(mel),
                                                        _lib.ptr(mel_post), _lib.ptr(ws), ws.numel(),
                                                        _lib.current_stream(dev)))
        return {'mel': mel, 'mel_post': mel_post, 'dur': dur_hat, 'pitch': pitch_hat, 'energy': energy_hat,
                'mel_len': total}

    def generate(self, x: torch.Tensor, alpha=1.0,
                 pitch_function: Callable[[torch.Tensor], torch.Tensor] = lambda x: x,
                 energy_function: Callable[[torch.Tensor], torch.Tensor] = lambda x: x,
                 mel_post_alloc=None, lengths=None) -> Dict[str, torch.Tensor]:
        self.eval()
        with torch.no_grad():
            x = self._check_tokens(x)
            h = self._get_handle(x.device)
            lib = _lib.lib()
            lens = None if lengths is None else self._check_lengths(x, lengths)
            # stage B's prenet depends on the tokens only: let stage A start it on a side stream (x is not
            # touched between the two calls)
            _lib.check(lib.ftb_ft_set_option(h, _lib.FTB_OPT_OVERLAP_PRENET, 1))
            try:
                dur_hat, pitch_hat, energy_hat = self.predict(x, alpha, lens)
                pitch_hat = pitch_function(pitch_hat)
                energy_hat = energy_function(energy_hat)
                return self.synthesize(x, dur_hat, pitch_hat, energy_hat, mel_post_alloc, lens)
            finally:
                _lib.check(lib.ftb_ft_set_option(h, _lib.FTB_OPT_OVERLAP_PRENET, 0))

    def generate_ragged(self, x: torch.Tensor, lengths, alpha=1.0,
                        pitch_function: Callable[[torch.Tensor], torch.Tensor] = lambda x: x,
                        energy_function: Callable[[torch.Tensor], torch.Tensor] = lambda x: x,
                        mel_post_alloc=None) -> Dict[str, torch.Tensor]:
        """The reference's per-sentence loop (gen_forward.py:106-118: one ``generate`` call per sentence, B = 1) as ONE
        padded batch: ``x`` (B, T) holds ``lengths[b]`` tokens per row (padding id arbitrary).  Row b of every output
        equals what ``generate(x[b:b+1, :lengths[b]])`` returns for that sentence alone, cut at ``mel_len[b]`` frames
        (``dur`` / ``pitch`` / ``energy`` at the first ``lengths[b]`` positions).  Plain ``generate`` on a padded batch
        keeps the reference's no-mask semantics instead: pad tokens are ordinary symbols there."""
        return self.generate(x, alpha, pitch_function, energy_function, mel_post_alloc, lengths=lengths)

    def generate_jit(self, x: torch.Tensor, alpha: float = 1.0, beta: float = 1.0) -> Dict[str, torch.Tensor]:
        """The TorchScript-exported entry point of the reference (models/forward_tacotron.py:270-284): ``generate``
        with the pitch scaled by ``beta`` and no callbacks.  (It does not switch to eval mode there either.)"""
        with torch.no_grad():
            dur_hat, pitch_hat, energy_hat = self.predict(x, alpha)
            return self.synthesize(x, dur_hat, pitch_hat * beta, energy_hat)

    def forward(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Teacher-forced forward in EVAL mode (models/forward_tacotron.py:184-242) -- the ground-truth-aligned
        feature dump of train_forward.py:33-52.  ``batch``: x (B,T) int64, mel (B,n_mels,Lm), mel_len (B,), dur (B,T),
        pitch (B,T), energy (B,T).  Durations / pitch / energy of the batch drive the synthesis, the decoder LSTM runs
        over packed sequences (rows stop at ``mel_len``, padded with ``padding_value``), the outputs are cut / padded
        to ``mel.size(2)``; 'dur' / 'pitch' / 'energy' are the raw predictor outputs (no fallback).  Training mode
        (dropout, BatchNorm statistics, autograd) is outside the path this package implements."""
        if self.training:
            raise NotImplementedError('training-mode forward() is outside the hot path this package implements '
                                      '(SURVEY 8f-4); call .eval() for the teacher-forced inference pass')
        with torch.no_grad():
            x = self._check_tokens(batch['x'])
            lib, dev = _lib.lib(), x.device
            B, T = x.shape
            mel_lens = batch['mel_len'].to(device=dev, dtype=torch.int32).contiguous()
            dur = batch['dur']
            if dur.dtype != torch.float32 or not dur.is_contiguous() or dur.device != dev:
                dur = dur.to(device=dev, dtype=torch.float32).contiguous()
            dur_hat = self.run_series_predictor('dur_pred', x).squeeze(-1)
            pitch_hat = self.run_series_predictor('pitch_pred', x).transpose(1, 2)
            energy_hat = self.run_series_predictor('energy_pred', x).transpose(1, 2)
            pitch = batch['pitch'].to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
            energy = batch['energy'].to(device=dev, dtype=torch.float32).reshape(B, T).contiguous()
            cum, total = LengthRegulator.plan(dur)           # clamps dur in place like the reference's lr()
            stats = torch.stack([total.max().to(torch.int64), mel_lens.max().to(torch.int64)]).tolist()  # one D2H
            L_lr, L = int(stats[0]), int(stats[1])
            if L <= 0 or L > L_lr:
                raise RuntimeError(f'mel_len (max {L}) must be positive and not exceed the expanded length {L_lr} '
                                   '(pack_padded_sequence would raise)')
            h = self._get_handle(dev)
            ws = self._workspace_for(h, B, T, L, dev)
            n_mels = self._dims['n_mels']
            mel = torch.empty((B, n_mels, L), dtype=torch.float32, device=dev)
            mel_post = torch.empty((B, n_mels, L), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.ftb_ft_synthesize_packed(h, _lib.ptr(x), _lib.ptr(cum), _lib.ptr(pitch), _lib.ptr(energy),
                                                        _lib.ptr(mel_lens), float(self.padding_value), B, T, L,
                                                        _lib.ptr(mel), _lib.ptr(mel_post), _lib.ptr(ws), ws.numel(),
                                                        _lib.current_stream(dev)))
            max_len = int(batch['mel'].size(2))
            return {'mel': self._pad(mel, max_len), 'mel_post': self._pad(mel_post, max_len),
                    'dur': dur_hat, 'pitch': pitch_hat, 'energy': energy_hat}

    def _pad(self, x: torch.Tensor, max_len: int) -> torch.Tensor:
        x = x[:, :, :max_len]
        return torch.nn.functional.pad(x, [0, max_len - x.size(2), 0, 0], 'constant', self.padding_value)

    def last_launch_count(self) -> int:
        return int(_lib.lib().ftb_ft_last_launch_count(self._handle)) if self._handle is not None else 0

    # ------------------------------------------------------------------ sub-module entry points
    def run_series_predictor(self, name: str, x: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
        """SeriesPredictor.forward(x, alpha) -> (B, T, 1)   (models/forward_tacotron.py:44-55)."""
        x = self._check_tokens(x)
        lib, dev = _lib.lib(), x.device
        B, T = x.shape
        h = self._get_handle(dev)
        ws = self._workspace_for(h, B, T, 0, dev)
        out = torch.empty((B, T, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ftb_ft_series_predictor(h, self._PREDICTORS[name], _lib.ptr(x), B, T, float(alpha),
                                                   _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.current_stream(dev)))
        return out

    def run_cbhg(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """CBHG.forward: x (B, C_in, S) -> (B, S, 2*channels)   (models/common_layers.py:86-119)."""
        _require_cuda(x, 'run_cbhg')
        lib, dev = _lib.lib(), x.device
        which = {'prenet': 0, 'postnet': 1}[name]
        B, Cn, S = x.shape
        h = self._get_handle(dev)
        xc = x.to(torch.float32).transpose(1, 2).contiguous()
        ch = self._dims['prenet_dims' if which == 0 else 'postnet_dims']
        ws = self._workspace_for(h, B, S, S, dev)
        out = torch.empty((B, S, 2 * ch), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ftb_ft_cbhg(h, which, _lib.ptr(xc), B, S, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                       _lib.current_stream(dev)))
        return out

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> 'ForwardTacotron':
        model_config = config['forward_tacotron']['model']
        model_config['num_chars'] = NUM_PHONEMES
        model_config['n_mels'] = config['dsp']['num_mels']
        return ForwardTacotron(**model_config)

    @classmethod
    def from_checkpoint(cls, path: Union[Path, str]) -> 'ForwardTacotron':
        checkpoint = torch.load(path, map_location=torch.device('cpu'))
        model = ForwardTacotron.from_config(checkpoint['config'])
        model.load_state_dict(checkpoint['model'])
        return model
