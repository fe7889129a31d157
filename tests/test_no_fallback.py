"""The product path must fail loudly without the GPU extension and must never touch the oracle."""
import re
from pathlib import Path

import pytest
import torch

from forwardtacotron_b200.models.common_layers import LengthRegulator
from forwardtacotron_b200.utils import synth
from forwardtacotron_b200.utils.dsp import DSP
from forwardtacotron_b200.utils.config import default_config

PKG = Path(__file__).resolve().parent.parent / 'forwardtacotron_b200'


def test_package_never_imports_oracle_or_reference():
    for py in PKG.rglob('*.py'):
        src = py.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), py
        assert '/root/reference' not in src, py


@pytest.mark.parametrize('kind', ['forward_tacotron', 'fast_pitch'])
def test_generate_on_cpu_raises(kind):
    model, _ = synth.synthetic_model(kind)
    x = synth.synthetic_tokens(1, 8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model.generate(x)


def test_length_regulator_on_cpu_raises():
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        LengthRegulator()(torch.zeros(1, 4, 8), torch.ones(1, 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_dsp_without_gpu_raises():
    dsp = DSP.from_config(default_config())
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        dsp.wav_to_mel(torch.zeros(2048).numpy())


def test_training_forward_is_out_of_scope():
    model, _ = synth.synthetic_model('forward_tacotron')
    with pytest.raises(NotImplementedError):
        model({'x': None})
