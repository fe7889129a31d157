#!/usr/bin/env python
"""Developer tool: where does the fused CBHG tail differ from the layer-by-layer path at the cfg2 shape?"""
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib
from forwardtacotron_b200.utils import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model, cfg = synth.synthetic_model('forward_tacotron')
model = model.cuda()
x = synth.synthetic_tokens(B, 200).cuda()
a = model.generate(x)
_lib.check(_lib.lib().ftb_ft_set_option(model._handle, _lib.FTB_OPT_UNFUSED_TAIL, 1))
b = model.generate(x)
torch.cuda.synchronize()
for k in ('mel', 'mel_post'):
    d = (a[k] - b[k]).abs()           # (B, 80, L)
    print(k, 'max', float(d.max()), 'mean', float(d.mean()))
d = (a['mel_post'] - b['mel_post']).abs().amax(dim=1)   # (B, L)
L = d.shape[1]
bad = (d > 0).nonzero()
print('L', L, 'bad frames', len(bad), 'of', d.numel())
if len(bad):
    rows = bad[:, 0] * L + bad[:, 1]
    tiles = torch.unique(rows // 128)
    print('bad 128-row tiles:', len(tiles), 'first', tiles[:40].tolist(), 'last', tiles[-10:].tolist())
    pairs = torch.unique(tiles // 2)
    print('bad pairs mod 74:', sorted(set((pairs % 74).tolist()))[:80])
    print('bad pairs // 74 (list position):', sorted(set((pairs // 74).tolist())))
