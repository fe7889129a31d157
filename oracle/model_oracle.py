"""fp32 CPU restatement of the reference's generate() path, written as pure
functions over a ``state_dict`` (TEST INFRASTRUCTURE ONLY, see oracle/__init__).

Every function cites the reference lines it restates (paths relative to the
upstream repo).  The restatement is checked against the imported reference by
``oracle/make_golden.py`` and by ``tests/test_oracle_golden.py`` against the
frozen fixtures.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
_IDENT = lambda t: t  # noqa: E731


# --------------------------------------------------------------------------
# shared building blocks
# --------------------------------------------------------------------------
def conv_relu_bn(sd: SD, p: str, x: torch.Tensor, relu: bool) -> torch.Tensor:
    """``BatchNormConv``: conv (no bias, pad k//2) -> optional ReLU -> eval BN.
    models/common_layers.py:38-52, models/forward_tacotron.py:58-71.  x: (B,C,S)."""
    w = sd[p + '.conv.weight']
    y = F.conv1d(x, w, None, 1, w.shape[2] // 2)
    if relu:
        y = torch.relu(y)
    inv = torch.rsqrt(sd[p + '.bnorm.running_var'] + 1e-5) * sd[p + '.bnorm.weight']
    sh = sd[p + '.bnorm.bias'] - sd[p + '.bnorm.running_mean'] * inv
    return y * inv[None, :, None] + sh[None, :, None]


def rnn_explicit(sd: SD, p: str, x: torch.Tensor, kind: str) -> torch.Tensor:
    """Bidirectional 1-layer GRU/LSTM written out step by step (numerics spec,
    SURVEY appendix A).  Zero initial state; the reverse direction runs over the
    padded sequence from t=S-1.  x: (B,S,I) -> (B,S,2H)."""
    outs = []
    for suffix, order in (('', range(x.shape[1])), ('_reverse', range(x.shape[1] - 1, -1, -1))):
        w_ih, w_hh = sd[f'{p}.weight_ih_l0{suffix}'], sd[f'{p}.weight_hh_l0{suffix}']
        b_ih, b_hh = sd[f'{p}.bias_ih_l0{suffix}'], sd[f'{p}.bias_hh_l0{suffix}']
        H = w_hh.shape[1]
        h = x.new_zeros(x.shape[0], H)
        c = x.new_zeros(x.shape[0], H)
        out = x.new_zeros(x.shape[0], x.shape[1], H)
        for t in order:
            gi = x[:, t] @ w_ih.T + b_ih
            gh = h @ w_hh.T + b_hh
            if kind == 'gru':  # gate order r, z, n
                r = torch.sigmoid(gi[:, :H] + gh[:, :H])
                z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                h = (1 - z) * n + z * h
            else:  # lstm, gate order i, f, g, o
                a = gi + gh
                i, f = torch.sigmoid(a[:, :H]), torch.sigmoid(a[:, H:2 * H])
                g, o = torch.tanh(a[:, 2 * H:3 * H]), torch.sigmoid(a[:, 3 * H:])
                c = f * c + i * g
                h = o * torch.tanh(c)
            out[:, t] = h
        outs.append(out)
    return torch.cat(outs, dim=2)


_RNN_CACHE: dict = {}


def rnn(sd: SD, p: str, x: torch.Tensor, kind: str) -> torch.Tensor:
    """Same maths as :func:`rnn_explicit` through ATen's fused CPU RNN (fast
    enough to serve as the CPU baseline).  models/common_layers.py:84,118;
    models/forward_tacotron.py:39,53,165-168,321."""
    w_hh = sd[f'{p}.weight_hh_l0']
    key = (id(w_hh), p, kind)
    hit = _RNN_CACHE.get(key)
    # The entry keeps the tensor it was built from: an id() alone is reused by Python once the old tensor is collected,
    # and a state_dict of the same shape would then silently get the previous weights (seen as a rare, exactly repeating
    # "kernel" mismatch in tests/test_gpu_rnn.py).
    if hit is None or hit[0] is not w_hh:
        cls = torch.nn.GRU if kind == 'gru' else torch.nn.LSTM
        mod = cls(sd[f'{p}.weight_ih_l0'].shape[1], w_hh.shape[1], batch_first=True, bidirectional=True)
        mod.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + '.')})
        mod.eval()
        if len(_RNN_CACHE) >= 64:
            _RNN_CACHE.pop(next(iter(_RNN_CACHE)))
        _RNN_CACHE[key] = hit = (w_hh, mod)
    with torch.no_grad():
        return hit[1](x)[0]


def length_regulate(x: torch.Tensor, dur: torch.Tensor) -> torch.Tensor:
    """models/common_layers.py:12-19.  Clamps ``dur`` IN PLACE at 0, rounds with
    trunc(dur + 0.5), repeats rows, zero-pads to the longest utterance."""
    dur.clamp_(min=0.)
    reps = (dur + 0.5).long()
    total = reps.sum(dim=1)
    L = int(total.max())
    out = x.new_zeros(x.shape[0], L, x.shape[2])
    for b in range(x.shape[0]):
        idx = torch.repeat_interleave(torch.arange(x.shape[1]), reps[b])
        out[b, :idx.numel()] = x[b, idx]
    return out


def apply_duration_fallback(dur: torch.Tensor) -> torch.Tensor:
    """models/forward_tacotron.py:254-255 / models/fast_pitch.py:295-296:
    batch-global test on trunc-toward-zero of the raw prediction."""
    if int(dur.long().sum()) <= 0:
        dur.fill_(2.)
    return dur


# --------------------------------------------------------------------------
# ForwardTacotron
# --------------------------------------------------------------------------
def ft_series_predictor(sd: SD, p: str, tokens: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """models/forward_tacotron.py:44-55: emb -> 3x(conv5, ReLU, BN) -> biGRU -> lin -> /alpha.
    tokens (B,T) int64 -> (B,T,1)."""
    x = sd[p + '.embedding.weight'][tokens].transpose(1, 2)
    for i in range(3):
        x = conv_relu_bn(sd, f'{p}.convs.{i}', x, relu=True)
    x = rnn(sd, p + '.rnn', x.transpose(1, 2), 'gru')
    x = x @ sd[p + '.lin.weight'].T + sd[p + '.lin.bias']
    return x / alpha


def highway(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """models/common_layers.py:30-35."""
    x1 = x @ sd[p + '.W1.weight'].T + sd[p + '.W1.bias']
    g = torch.sigmoid(x @ sd[p + '.W2.weight'].T + sd[p + '.W2.bias'])
    return g * torch.relu(x1) + (1. - g) * x


def cbhg(sd: SD, p: str, x: torch.Tensor, return_stages: bool = False):
    """models/common_layers.py:86-119.  x: (B,C,S) -> (B,S,2*channels)."""
    S = x.shape[-1]
    n_bank = len([k for k in sd if k.startswith(p + '.conv1d_bank.') and k.endswith('.conv.weight')])
    n_hw = len([k for k in sd if k.startswith(p + '.highways.') and k.endswith('.W1.weight')])
    bank = torch.cat([conv_relu_bn(sd, f'{p}.conv1d_bank.{i}', x, True)[:, :, :S] for i in range(n_bank)], dim=1)
    pooled = F.max_pool1d(bank, 2, 1, 1)[:, :, :S]
    p1 = conv_relu_bn(sd, p + '.conv_project1', pooled, True)
    p2 = conv_relu_bn(sd, p + '.conv_project2', p1, False) + x
    h = p2.transpose(1, 2) @ sd[p + '.pre_highway.weight'].T
    for i in range(n_hw):
        h = highway(sd, f'{p}.highways.{i}', h)
    out = rnn(sd, p + '.rnn', h, 'gru')
    if return_stages:
        return out, {'bank_pooled': pooled, 'proj1': p1, 'proj2_res': p2, 'highway': h}
    return out


def cond_proj(sd: SD, p: str, series: torch.Tensor) -> torch.Tensor:
    """Conv1d(1->C, k3, pad1, bias) on (B,1,T), returned channel-last (B,T,C).
    models/forward_tacotron.py:145-146,308-314."""
    return F.conv1d(series, sd[p + '.weight'], sd[p + '.bias'], 1, 1).transpose(1, 2)


def ft_predict(sd: SD, tokens: torch.Tensor, alpha: float = 1.0):
    """Stage A of generate (models/forward_tacotron.py:251-262 before the callbacks)."""
    dur = apply_duration_fallback(ft_series_predictor(sd, 'dur_pred', tokens, alpha).squeeze(2))
    pitch = ft_series_predictor(sd, 'pitch_pred', tokens).transpose(1, 2)
    energy = ft_series_predictor(sd, 'energy_pred', tokens).transpose(1, 2)
    return dur, pitch, energy


def ft_synthesize(sd: SD, tokens, dur, pitch, energy, pitch_strength=1.0, energy_strength=1.0,
                  return_stages: bool = False) -> Dict[str, torch.Tensor]:
    """models/forward_tacotron.py:289-330 (_generate_mel).  ``dur`` is clamped in place."""
    x = sd['embedding.weight'][tokens].transpose(1, 2)
    enc = cbhg(sd, 'prenet', x)
    enc = enc + cond_proj(sd, 'pitch_proj', pitch) * pitch_strength
    enc = enc + cond_proj(sd, 'energy_proj', energy) * energy_strength
    up = length_regulate(enc, dur)
    dec = rnn(sd, 'lstm', up, 'lstm')
    mel = (dec @ sd['lin.weight'].T + sd['lin.bias']).transpose(1, 2)
    post = cbhg(sd, 'postnet', mel)
    mel_post = (post @ sd['post_proj.weight'].T).transpose(1, 2)
    out = {'mel': mel, 'mel_post': mel_post, 'dur': dur, 'pitch': pitch, 'energy': energy}
    if return_stages:
        out['_enc'] = enc
        out['_up'] = up
        out['_dec'] = dec
    return out


def ft_generate(sd: SD, tokens: torch.Tensor, alpha: float = 1.0,
                pitch_function: Callable = _IDENT, energy_function: Callable = _IDENT,
                pitch_strength: float = 1.0, energy_strength: float = 1.0) -> Dict[str, torch.Tensor]:
    """models/forward_tacotron.py:244-268."""
    with torch.no_grad():
        dur, pitch, energy = ft_predict(sd, tokens, alpha)
        return ft_synthesize(sd, tokens, dur, pitch_function(pitch), energy_function(energy),
                             pitch_strength, energy_strength)


def ft_forward(sd: SD, batch: Dict[str, torch.Tensor], pitch_strength: float = 1.0, energy_strength: float = 1.0,
               padding_value: float = -11.5129) -> Dict[str, torch.Tensor]:
    """Teacher-forced ``forward`` in eval mode, models/forward_tacotron.py:184-242 (the GTA dump of
    train_forward.py:33-52).  The packed-sequence LSTM (:224-230) is restated row by row: row b runs the bidirectional
    LSTM over its first ``mel_len[b]`` frames only, the rest of the row is ``padding_value``; the result is cut to
    ``max(mel_len)`` frames (pad_packed_sequence), ``lin`` / postnet / ``post_proj`` see the padded rows, and both
    outputs are cut / padded to ``mel.size(2)`` (:238-239, ``_pad`` :332-335).  ``batch['dur']`` is clamped in place."""
    with torch.no_grad():
        x, mel_lens = batch['x'], batch['mel_len']
        dur_hat = ft_series_predictor(sd, 'dur_pred', x).squeeze(-1)
        pitch_hat = ft_series_predictor(sd, 'pitch_pred', x).transpose(1, 2)
        energy_hat = ft_series_predictor(sd, 'energy_pred', x).transpose(1, 2)
        enc = cbhg(sd, 'prenet', sd['embedding.weight'][x].transpose(1, 2))
        enc = enc + cond_proj(sd, 'pitch_proj', batch['pitch'].unsqueeze(1)) * pitch_strength
        enc = enc + cond_proj(sd, 'energy_proj', batch['energy'].unsqueeze(1)) * energy_strength
        up = length_regulate(enc, batch['dur'])
        L = int(mel_lens.max())
        dec = up.new_full((up.shape[0], L, 2 * sd['lstm.weight_hh_l0'].shape[1]), padding_value)
        for b in range(up.shape[0]):
            n = int(mel_lens[b])
            dec[b, :n] = rnn(sd, 'lstm', up[b:b + 1, :n], 'lstm')[0]
        mel = (dec @ sd['lin.weight'].T + sd['lin.bias']).transpose(1, 2)
        post = cbhg(sd, 'postnet', mel)
        mel_post = (post @ sd['post_proj.weight'].T).transpose(1, 2)
        max_len = batch['mel'].size(2)

        def pad(t):
            t = t[:, :, :max_len]
            return F.pad(t, [0, max_len - t.size(2), 0, 0], 'constant', padding_value)
        return {'mel': pad(mel), 'mel_post': pad(mel_post), 'dur': dur_hat, 'pitch': pitch_hat, 'energy': energy_hat}


# --------------------------------------------------------------------------
# FastPitch
# --------------------------------------------------------------------------
def layer_norm(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def mha(sd: SD, p: str, x: torch.Tensor, heads: int, key_pad: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.MultiheadAttention self-attention, batch-first restatement
    (models/fast_pitch.py:64,80-82).  x: (B,S,E); key_pad: (B,S) bool, True = ignore."""
    B, S, E = x.shape
    hd = E // heads
    qkv = x @ sd[p + '.in_proj_weight'].T + sd[p + '.in_proj_bias']
    q, k, v = (t.reshape(B, S, heads, hd).transpose(1, 2) for t in qkv.split(E, dim=2))
    att = (q * (1.0 / math.sqrt(hd))) @ k.transpose(2, 3)
    if key_pad is not None:
        att = att.masked_fill(key_pad[:, None, None, :], float('-inf'))
    ctx = (torch.softmax(att, dim=-1) @ v).transpose(1, 2).reshape(B, S, E)
    return ctx @ sd[p + '.out_proj.weight'].T + sd[p + '.out_proj.bias']


def fft_block(sd: SD, p: str, x: torch.Tensor, heads: int, key_pad) -> torch.Tensor:
    """models/fast_pitch.py:76-92 (post-LN; conv1 k9 pad4 + ReLU; conv2 k1)."""
    x = layer_norm(sd, p + '.norm1', x + mha(sd, p + '.self_attn', x, heads, key_pad))
    w1, w2 = sd[p + '.conv1.weight'], sd[p + '.conv2.weight']
    f = torch.relu(F.conv1d(x.transpose(1, 2), w1, sd[p + '.conv1.bias'], 1, w1.shape[2] // 2))
    f = F.conv1d(f, w2, sd[p + '.conv2.bias'], 1, w2.shape[2] // 2).transpose(1, 2)
    return layer_norm(sd, p + '.norm2', x + f)


def forward_transformer(sd: SD, p: str, x: torch.Tensor, heads: int, key_pad=None) -> torch.Tensor:
    """models/fast_pitch.py:121-130.  x: (B,S,E).  ``pe`` and ``scale`` are read
    from the state_dict; S > 5000 raises like the reference's broadcast does."""
    S = x.shape[1]
    pe = sd[p + '.pos_encoder.pe']
    if S > pe.shape[0]:
        raise RuntimeError(f'The size of tensor a ({S}) must match the size of tensor b ({pe.shape[0]}) '
                           'at non-singleton dimension 0')
    x = x + sd[p + '.pos_encoder.scale'] * pe[:S, 0][None]
    n_layers = len([k for k in sd if k.startswith(p + '.layers.') and k.endswith('.norm1.weight')])
    for i in range(n_layers):
        x = fft_block(sd, f'{p}.layers.{i}', x, heads, key_pad)
    return layer_norm(sd, p + '.norm', x)


def fp_series_predictor(sd: SD, p: str, tokens, heads: int, alpha: float = 1.0) -> torch.Tensor:
    """models/fast_pitch.py:152-160 (called without a padding mask from generate)."""
    x = forward_transformer(sd, p + '.transformer', sd[p + '.embedding.weight'][tokens], heads)
    return (x @ sd[p + '.lin.weight'].T + sd[p + '.lin.bias']) / alpha


def fp_predict(sd: SD, tokens, alpha: float = 1.0, heads=(2, 2, 2)):
    dur = apply_duration_fallback(fp_series_predictor(sd, 'dur_pred', tokens, heads[0], alpha).squeeze(2))
    pitch = fp_series_predictor(sd, 'pitch_pred', tokens, heads[1]).transpose(1, 2)
    energy = fp_series_predictor(sd, 'energy_pred', tokens, heads[2]).transpose(1, 2)
    return dur, pitch, energy


def fp_synthesize(sd: SD, tokens, dur, pitch, energy, pitch_strength=1.0, energy_strength=1.0,
                  prenet_heads: int = 2, postnet_heads: int = 2) -> Dict[str, torch.Tensor]:
    """models/fast_pitch.py:313-340: key-padding mask (tokens == 0) on the prenet
    only; postnet unmasked; 'mel_post' is the same tensor as 'mel'."""
    x = forward_transformer(sd, 'prenet', sd['embedding.weight'][tokens], prenet_heads, key_pad=(tokens == 0))
    x = x + cond_proj(sd, 'pitch_proj', pitch) * pitch_strength
    x = x + cond_proj(sd, 'energy_proj', energy) * energy_strength
    x = length_regulate(x, dur)
    x = forward_transformer(sd, 'postnet', x, postnet_heads)
    mel = (x @ sd['lin.weight'].T + sd['lin.bias']).transpose(1, 2)
    return {'mel': mel, 'mel_post': mel, 'dur': dur, 'pitch': pitch, 'energy': energy}


def fp_forward(sd: SD, batch: Dict[str, torch.Tensor], pitch_strength=1.0, energy_strength=1.0,
               heads=(2, 2, 2, 2, 2), padding_value: float = -11.5129) -> Dict[str, torch.Tensor]:
    """Teacher-forced ``FastPitch.forward`` in eval mode, models/fast_pitch.py:243-283: predictors and prenet with the
    token padding mask (:255-261), batch durations / pitch / energy, postnet with the mel-length key mask (:274-278),
    outputs cut / padded to ``mel.size(2)`` (:283-284).  ``batch['dur']`` is clamped in place."""
    with torch.no_grad():
        x, mel_lens = batch['x'], batch['mel_len']
        tok_mask = x == 0

        def series(p, hd):
            t = forward_transformer(sd, p + '.transformer', sd[p + '.embedding.weight'][x], hd, key_pad=tok_mask)
            return t @ sd[p + '.lin.weight'].T + sd[p + '.lin.bias']
        dur_hat = series('dur_pred', heads[0]).squeeze(-1)
        pitch_hat = series('pitch_pred', heads[1]).transpose(1, 2)
        energy_hat = series('energy_pred', heads[2]).transpose(1, 2)
        h = forward_transformer(sd, 'prenet', sd['embedding.weight'][x], heads[3], key_pad=tok_mask)
        h = h + cond_proj(sd, 'pitch_proj', batch['pitch'].unsqueeze(1)) * pitch_strength
        h = h + cond_proj(sd, 'energy_proj', batch['energy'].unsqueeze(1)) * energy_strength
        h = length_regulate(h, batch['dur'])
        len_mask = torch.arange(h.shape[1])[None, :] >= mel_lens[:, None]
        h = forward_transformer(sd, 'postnet', h, heads[4], key_pad=len_mask)
        mel = (h @ sd['lin.weight'].T + sd['lin.bias']).transpose(1, 2)
        max_len = batch['mel'].size(2)
        mel = F.pad(mel[:, :, :max_len], [0, max_len - min(max_len, mel.size(2)), 0, 0], 'constant', padding_value)
        return {'mel': mel, 'mel_post': mel, 'dur': dur_hat, 'pitch': pitch_hat, 'energy': energy_hat}


def fp_generate(sd: SD, tokens, alpha: float = 1.0, pitch_function: Callable = _IDENT,
                energy_function: Callable = _IDENT, pitch_strength=1.0, energy_strength=1.0,
                heads=(2, 2, 2, 2, 2)) -> Dict[str, torch.Tensor]:
    """models/fast_pitch.py:286-303.  heads = (dur, pitch, energy, prenet, postnet)."""
    with torch.no_grad():
        dur, pitch, energy = fp_predict(sd, tokens, alpha, heads[:3])
        return fp_synthesize(sd, tokens, dur, pitch_function(pitch), energy_function(energy),
                             pitch_strength, energy_strength, heads[3], heads[4])
