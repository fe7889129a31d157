"""Seeded synthetic weights and inputs: there are no checkpoints or datasets offline.

Plain random init is useless for this path: the raw duration prediction is ~0, so ``generate``
takes the ``fill_(2.0)`` fallback (models/forward_tacotron.py:254-255) and BatchNorm is the
identity.  The recipe below keeps the reference's state_dict layout and
 * randomises every BatchNorm's running stats / affine (exercises the ReLU->BN epilogues),
 * gives the highway ``W1.bias`` (zero at init) a non-zero value,
 * rescales ``dur_pred.lin`` so rounded durations average ~6 frames with a real spread.
The constants were calibrated once with the CPU oracle (oracle/make_golden.py --calibrate).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .config import default_config

# (scale, bias) applied to dur_pred.lin for seed 0 so that durations are ~6 +- 1.5 frames
DUR_CALIBRATION = {
    'forward_tacotron': (60.0, -7.9),
    'fast_pitch': (3.0, 4.4),
}


def synthetic_model(tts_model: str = 'forward_tacotron', config: Optional[Dict[str, Any]] = None, seed: int = 0,
                    dur_scale: Optional[float] = None, dur_bias: Optional[float] = None, mel_gain: float = 1.0,
                    plain_init: bool = False):
    """Build a model on the CPU with the synthetic-weight recipe.  ``plain_init`` keeps the untouched
    random init (covers the duration fallback branch).  ``mel_gain`` scales the output heads so mels
    reach trained-checkpoint magnitude (the stress case for the absolute tolerance)."""
    from .checkpoints import init_tts_model
    config = config or default_config(tts_model)
    config['tts_model'] = tts_model
    torch.manual_seed(seed)
    model = init_tts_model(config)
    if plain_init:
        return model, config
    g = torch.Generator().manual_seed(123 + seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        for name, p in model.named_parameters():
            if name.endswith('.W1.bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        s, b = DUR_CALIBRATION[tts_model]
        s = dur_scale if dur_scale is not None else s
        b = dur_bias if dur_bias is not None else b
        model.dur_pred.lin.weight.mul_(s)
        model.dur_pred.lin.bias.fill_(b)
        if mel_gain != 1.0:
            model.lin.weight.mul_(mel_gain)
            model.lin.bias.mul_(mel_gain)
            if hasattr(model, 'post_proj'):
                model.post_proj.weight.mul_(mel_gain)
    return model, config


def synthetic_tokens(B: int, T: int, seed: int = 1, ragged: bool = False) -> torch.Tensor:
    """Phoneme ids in [1, 135) (no pad id 0); ``ragged`` zero-pads rows to random lengths >= T//2 to
    exercise the reference's no-mask semantics and the FastPitch key-padding mask."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(1, 135, (B, T), generator=g, dtype=torch.long)
    if ragged:
        lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        lens[0] = T
        x = x * (torch.arange(T)[None, :] < lens[:, None])
    return x


def synthetic_audio(n_clips: int, seed: int = 7, min_s: float = 2.0, max_s: float = 10.0, sr: int = 22050):
    """cfg4 clips: durations U(min_s, max_s); 0.1*N(0,1) noise, every 16th clip a sine, every 17th silence
    (exercises the 1e-5 clamp).  Returns (flat float32 audio, int64 clip offsets)."""
    g = torch.Generator().manual_seed(seed)
    lens = (torch.rand(n_clips, generator=g) * (max_s - min_s) + min_s).mul(sr).long()
    offs = torch.zeros(n_clips + 1, dtype=torch.long)
    offs[1:] = torch.cumsum(lens, 0)
    audio = torch.randn(int(offs[-1]), generator=g) * 0.1
    for i in range(n_clips):
        a, b = int(offs[i]), int(offs[i + 1])
        if i % 16 == 5:
            t = torch.arange(b - a, dtype=torch.float32) / sr
            audio[a:b] = 0.5 * torch.sin(2 * torch.pi * (220.0 + 10 * i) * t)
        elif i % 17 == 9:
            audio[a:b] = 0.0
    return audio, offs
