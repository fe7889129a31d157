import sys, time
sys.path.insert(0, '.')
import torch
from forwardtacotron_b200.utils import synth
dev = torch.device('cuda', 0)
model, _ = synth.synthetic_model('fast_pitch')
model = model.to(dev)
K = 9
for S in (1, 2, 3):
    xs = [synth.synthetic_tokens(128, 300, seed=5 + i).to(dev) for i in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    for i in range(S):
        with torch.cuda.stream(streams[i]):
            for _ in range(2):
                out = model.generate(xs[i])
    torch.cuda.synchronize()
    frames = int(out['mel_len'].sum())
    t0 = time.perf_counter()
    for k in range(K):
        with torch.cuda.stream(streams[k % S]):
            out = model.generate(xs[k % S])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'fast_pitch lanes {S}: {dt / K * 1e3:.2f} ms per generate, {frames * K / dt / 1e6:.2f} M frames/s')
