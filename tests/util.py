"""Shared helpers of the test-suite (the oracle is imported HERE, never by the package)."""
from pathlib import Path

import numpy as np
import torch

from forwardtacotron_b200.utils import synth

GOLD = Path(__file__).resolve().parent / 'golden'

# north-star tolerance for floating-point outputs (BASELINE.json): max-abs 1e-2, mean-abs 1e-3
MAX_ABS, MEAN_ABS = 1e-2, 1e-3


def load(name):
    return {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(GOLD / f'{name}.npz').items()}


def cuda_model(kind, gemm_mode=0, **kw):
    model, cfg = synth.synthetic_model(kind, **kw)
    model.gemm_mode = gemm_mode
    return model.cuda(), cfg


def cpu_state_dict(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def err(a, b):
    d = (a.detach().cpu().float() - b.detach().cpu().float()).abs()
    return float(d.max()) if d.numel() else 0.0, float(d.mean()) if d.numel() else 0.0


def assert_close(a, b, max_abs=MAX_ABS, mean_abs=MEAN_ABS, what=''):
    assert tuple(a.shape) == tuple(b.shape), f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    mx, mn = err(a, b)
    assert mx <= max_abs and mn <= mean_abs, f'{what}: max-abs {mx:.3e} (<= {max_abs}), mean-abs {mn:.3e} (<= {mean_abs})'
    return mx, mn


def rounded(dur):
    return (dur.detach().cpu().clamp(min=0) + 0.5).long()


def near_tie_mask(dur, eps=2e-4):
    """Durations whose rounding could legitimately flip under a different fp32 summation order."""
    frac = (dur.detach().cpu().clamp(min=0) + 0.5) % 1.0
    return (frac < eps) | (frac > 1 - eps)
