"""Row a9/a11 of the scope table: constructor / state_dict / checkpoint contract of the mirrored models."""
import copy

import pytest
import torch

from forwardtacotron_b200.models.fast_pitch import FastPitch
from forwardtacotron_b200.models.forward_tacotron import ForwardTacotron
from forwardtacotron_b200.utils import checkpoints, synth
from forwardtacotron_b200.utils.config import default_config


def test_forward_tacotron_layout():
    m = ForwardTacotron.from_config(default_config())
    sd = m.state_dict()
    assert len(sd) == 322                                    # SURVEY a9 [probe]
    assert sum(p.numel() for p in m.parameters()) == 24_509_235
    assert sd['step'].shape == (1,) and sd['step'].dtype == torch.int64
    assert sum(k.endswith('num_batches_tracked') for k in sd) == 37
    assert sd['prenet.conv1d_bank.15.conv.weight'].shape == (256, 256, 16)
    assert sd['prenet.conv_project1.conv.weight'].shape == (256, 4096, 3)
    assert sd['postnet.conv_project2.conv.weight'].shape == (80, 256, 3)
    assert sd['postnet.pre_highway.weight'].shape == (256, 80)
    assert sd['lstm.weight_hh_l0_reverse'].shape == (2048, 512)
    assert sd['pitch_pred.rnn.weight_hh_l0'].shape == (384, 128)
    assert sd['pitch_proj.weight'].shape == (512, 1, 3)
    assert sd['post_proj.weight'].shape == (80, 512) and 'post_proj.bias' not in sd
    assert m.get_step() == 0


def test_fast_pitch_layout():
    m = FastPitch.from_config(default_config('fast_pitch'))
    sd = m.state_dict()
    assert len(sd) == 277                                    # SURVEY a11 [probe]
    assert sd['prenet.pos_encoder.pe'].shape == (5000, 1, 256)
    assert sd['dur_pred.transformer.layers.3.self_attn.in_proj_weight'].shape == (384, 128)
    assert sd['postnet.layers.0.conv1.weight'].shape == (1024, 256, 9)
    assert sd['postnet.layers.0.conv2.weight'].shape == (256, 1024, 1)
    assert sd['lin.weight'].shape == (80, 256)


@pytest.mark.parametrize('kind,cls', [('forward_tacotron', ForwardTacotron), ('fast_pitch', FastPitch)])
def test_checkpoint_roundtrip(tmp_path, kind, cls):
    model, cfg = synth.synthetic_model(kind)
    path = tmp_path / 'latest_model.pt'
    checkpoints.save_checkpoint(model, None, copy.deepcopy(cfg), path)      # utils/checkpoints.py:16-18 format
    ck = torch.load(path, map_location='cpu')
    assert set(ck) == {'model', 'optim', 'config'}
    loaded = cls.from_checkpoint(path)
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), loaded.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    via_dispatch, _ = checkpoints.load_tts_model(path)
    assert type(via_dispatch) is cls
    with pytest.raises(ValueError):
        checkpoints.init_tts_model({'tts_model': 'tacotron'})


def test_synthetic_recipe_is_deterministic():
    a, _ = synth.synthetic_model('forward_tacotron')
    b, _ = synth.synthetic_model('forward_tacotron')
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    assert torch.equal(synth.synthetic_tokens(4, 9), synth.synthetic_tokens(4, 9))
    assert int(synth.synthetic_tokens(4, 9).min()) >= 1
    assert int((synth.synthetic_tokens(4, 20, ragged=True) == 0).sum()) > 0
