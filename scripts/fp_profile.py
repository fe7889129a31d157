#!/usr/bin/env python
"""Developer tool: per-kernel-family CUDA-event times of one FastPitch.generate (cfg3 shape)."""
import sys
sys.path.insert(0, '.')
import torch
import bench
from forwardtacotron_b200 import _lib
from forwardtacotron_b200.utils import synth

lib = _lib.lib()
dev = torch.device('cuda', 0)
model, _ = synth.synthetic_model('fast_pitch')
model = model.to(dev)
x = synth.synthetic_tokens(128, 300, seed=5).to(dev)
for _ in range(2):
    out = model.generate(x)
torch.cuda.synchronize()
lib.ftb_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = model.generate(x)
e1.record()
torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1), 'L', out['mel'].shape[-1])
for f in bench.collect_profile(lib, 1):
    tf = f['flops_per_step'] / f['ms_per_step'] / 1e9 if f['flops_per_step'] else 0
    print(f"{f['name']:22s} {f['ms_per_step']:8.3f} ms  x{f['launches_per_step']:.0f}  {tf:.0f} TF/s")
