#!/usr/bin/env python
"""Small end-to-end pass over every public entry point at tiny sizes (ForwardTacotron / FastPitch generate and
teacher-forced forward, wav_to_mel): the thing to run under a memory checker where one is available."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from forwardtacotron_b200.utils import synth  # noqa: E402
from forwardtacotron_b200.utils.config import default_config  # noqa: E402
from forwardtacotron_b200.utils.dsp import DSP  # noqa: E402

model, _ = synth.synthetic_model('forward_tacotron')
model = model.cuda().eval()
x = synth.synthetic_tokens(3, 20, seed=2).cuda()
out = model.generate(x)
B, T = x.shape
dur = torch.randint(1, 5, (B, T)).float().cuda()
batch = {'x': x, 'dur': dur, 'mel_len': (dur + 0.5).long().sum(1), 'pitch': torch.randn(B, T).cuda(),
         'energy': torch.randn(B, T).cuda(), 'mel': torch.zeros(B, 80, int(dur.sum(1).max()) + 2).cuda()}
fwd = model(batch)
fp, _ = synth.synthetic_model('fast_pitch')
fp = fp.cuda().eval()
o2 = fp.generate(x)
f2 = fp(batch)
mel = DSP.from_config(default_config()).wav_to_mel((0.1 * np.random.default_rng(0).standard_normal(9000)).astype(np.float32))
torch.cuda.synchronize()
print('ok', tuple(out['mel'].shape), tuple(fwd['mel'].shape), tuple(o2['mel'].shape), tuple(f2['mel'].shape), mel.shape)
