#!/usr/bin/env python
"""Developer tool (needs a build with FTB_NVCC_DEFINES=FTB_PHASE_TIMING: rm -rf forwardtacotron_b200/csrc/build && FTB_NVCC_DEFINES=FTB_PHASE_TIMING python __graft_entry__.py build): per-unit phase clocks of the fused CBHG tail kernel (CTA 0, first 128 accumulator units of the
postnet launch of one cfg2-sized generate()).  Unit = 256 accumulator columns of one layer of one 128-row tile:
1 (pre_highway) + 4 x 2 (highways) + 6 (GRU input projection) = 15 units of 256 columns per tile."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from forwardtacotron_b200 import _lib
from forwardtacotron_b200.utils import synth

fn = C.CDLL(str(_lib.lib_path())).ftb_debug_tail_timing
fn.argtypes = [C.c_void_p]
model, cfg = synth.synthetic_model('forward_tacotron')
model = model.cuda()
x = synth.synthetic_tokens(64, 200).cuda()
model.generate(x)
dbg = torch.zeros(128 * 8, dtype=torch.int64, device='cuda')
fn(dbg.data_ptr())
model.generate(x)
torch.cuda.synchronize()
fn(None)
d = dbg.cpu().view(128, 8)
t0 = int(d[0, 0])
print('unit  layer | mma: wait-start  w-landed  issued | epi: acc-seen  done | prod: loads issued     (clocks since unit 0)')
# unit sequence (csrc/cbhg_tail.cu, tail_period): slot X runs pre + highways with slot Y's projection dealt out in
# the gaps, then the roles swap; 256 accumulator columns per unit
def period(k, nx=9, ny=9):
    out = []
    for ph in (0, 1):
        s, o = 'XY'[ph], 'YX'[ph]
        tko = k - 1 if ph == 0 else k
        vo = tko >= 0
        out += [f'pre {s}{k}']
        for j in range(4):
            out += [f'hw{j} {s}{k}'] * 2
            if vo and j < 3:
                out += [f'in  {o}{tko}'] * 2
    return out
seq = period(0) + period(1) + period(2) + period(3)
for g in range(min(len(seq), 128)):
    r = [int(d[g, k]) - t0 if int(d[g, k]) else -1 for k in range(6)]
    print(f'{g:4d}  {seq[g]} | {r[0]:10d} {r[1]:10d} {r[2]:10d} | {r[3]:10d} {r[4]:10d} | {r[5]:10d}')
p0 = len(period(0))
UPR = len(period(1))
per_tile = (int(d[p0 + 2 * UPR, 0]) - int(d[p0 + UPR, 0])) // 2
print(f'clocks per tile (2nd tile): {per_tile}')
