"""Pins oracle.model_oracle.ft_forward / fp_forward (teacher-forced forward in eval mode, the GTA feature path) against
the REAL reference and freezes tests/golden/ft_forward_b3_t30.npz and fp_forward_b3_t30.npz.  Runs in the build container
only (/root/reference).

    python oracle/make_golden_forward.py
"""
from __future__ import annotations

import copy
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path('/root/reference')
GOLD = ROOT / 'tests' / 'golden'

from forwardtacotron_b200.utils import synth  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402


def synthetic_batch(B=3, T=30, seed=21):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(1, 135, (B, T), generator=g)
    dur = torch.randint(0, 9, (B, T), generator=g).float()
    dur[1, T // 2:] = 0.0          # a short utterance: its packed sequence ends well before the padded length
    dur[2, :3] = torch.tensor([-1.0, 2.5, 0.49])  # clamp / rounding edge cases of the LengthRegulator
    mel_len = (dur.clamp(min=0) + 0.5).long().sum(1)
    Lm = int(mel_len.max()) + 3    # the target mel is longer than the synthesised one: _pad appends padding_value
    return {'x': x, 'dur': dur, 'mel_len': mel_len, 'mel': torch.zeros(B, 80, Lm),
            'pitch': torch.randn(B, T, generator=g), 'energy': torch.randn(B, T, generator=g)}


def main():
    if not REF.exists():
        raise SystemExit('/root/reference is not present: golden generation runs in the build container only')
    sys.path.insert(0, str(REF))
    from models.forward_tacotron import ForwardTacotron as RefFT  # type: ignore
    torch.set_num_threads(8)
    model, cfg = synth.synthetic_model('forward_tacotron')
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref = RefFT.from_config(copy.deepcopy(cfg))
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    batch = synthetic_batch()
    with torch.no_grad():
        r = ref({k: v.clone() for k, v in batch.items()})
    o = mo.ft_forward(sd, {k: v.clone() for k, v in batch.items()}, ref.pitch_strength, ref.energy_strength,
                      ref.padding_value)
    errs = {k: float((r[k] - o[k]).abs().max()) for k in ('mel', 'mel_post', 'dur', 'pitch', 'energy')}
    assert r['mel'].shape == o['mel'].shape and max(errs.values()) < 2e-5, errs
    # cut shorter than the synthesis as well (mel.size(2) < max(mel_len)): _pad truncates
    b2 = {k: v.clone() for k, v in batch.items()}
    b2['mel'] = torch.zeros(3, 80, int(batch['mel_len'].max()) - 5)
    with torch.no_grad():
        r2 = ref({k: v.clone() for k, v in b2.items()})
    o2 = mo.ft_forward(sd, {k: v.clone() for k, v in b2.items()}, ref.pitch_strength, ref.energy_strength, ref.padding_value)
    assert r2['mel'].shape == o2['mel'].shape and float((r2['mel_post'] - o2['mel_post']).abs().max()) < 2e-5
    np.savez_compressed(GOLD / 'ft_forward_b3_t30.npz', x=batch['x'].numpy(), dur_in=batch['dur'].numpy(),
                        mel_len=batch['mel_len'].numpy(), mel_frames=np.int64(batch['mel'].size(2)),
                        pitch_in=batch['pitch'].numpy(), energy_in=batch['energy'].numpy(),
                        mel=r['mel'].numpy(), mel_post=r['mel_post'].numpy(), dur=r['dur'].numpy(),
                        pitch=r['pitch'].numpy(), energy=r['energy'].numpy())
    print('ft_forward_b3_t30', errs, 'mel', tuple(r['mel'].shape), 'mel_len', batch['mel_len'].tolist())

    # ---- FastPitch: token padding mask on predictors / prenet, mel-length key mask on the postnet
    from models.fast_pitch import FastPitch as RefFP  # type: ignore
    model, cfg = synth.synthetic_model('fast_pitch')
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref = RefFP.from_config(copy.deepcopy(cfg))
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    batch = synthetic_batch(seed=22)
    batch['x'][1, 30 // 2:] = 0        # padded tokens (their durations are already 0): exercises the token mask
    with torch.no_grad():
        r = ref({k: v.clone() for k, v in batch.items()})
    o = mo.fp_forward(sd, {k: v.clone() for k, v in batch.items()}, ref.pitch_strength, ref.energy_strength)
    errs = {k: float((r[k] - o[k]).abs().max()) for k in ('mel', 'mel_post', 'dur', 'pitch', 'energy')}
    assert r['mel'].shape == o['mel'].shape and max(errs.values()) < 2e-5, errs
    np.savez_compressed(GOLD / 'fp_forward_b3_t30.npz', x=batch['x'].numpy(), dur_in=batch['dur'].numpy(),
                        mel_len=batch['mel_len'].numpy(), mel_frames=np.int64(batch['mel'].size(2)),
                        pitch_in=batch['pitch'].numpy(), energy_in=batch['energy'].numpy(),
                        mel=r['mel'].numpy(), mel_post=r['mel_post'].numpy(), dur=r['dur'].numpy(),
                        pitch=r['pitch'].numpy(), energy=r['energy'].numpy())
    print('fp_forward_b3_t30', errs, 'mel', tuple(r['mel'].shape), 'mel_len', batch['mel_len'].tolist())


if __name__ == '__main__':
    main()
