#!/usr/bin/env python
"""2+ GPU check of the fused final gather (utils/peer_window.py): every rank's post_proj epilogue stores mel_post
straight into rank 0's memory over NVLink; the result must equal the NCCL gather bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/p2p_gather_check.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from forwardtacotron_b200.utils import batching, synth
from forwardtacotron_b200.utils.peer_window import PeerWindow


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    model, _ = synth.synthetic_model('forward_tacotron')
    model = model.to(dev)
    g = torch.Generator().manual_seed(3)
    n_utts = int(os.environ.get('N_UTTS', '192'))
    utts = [torch.randint(1, 135, (int(n),), generator=g).tolist() for n in torch.randint(40, 200, (n_utts,), generator=g)]
    window = PeerWindow(2 << 30)
    res = {}
    for mode in ('nccl', 'p2p', 'nccl', 'p2p'):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        out = batching.synthesize_corpus(model, utts, max_tokens=8192, window=window if mode == 'p2p' else None)
        torch.cuda.synchronize()
        dist.barrier()
        res[mode] = (time.perf_counter() - t0, out)
    if rank == 0:
        a, b = res['nccl'][1], res['p2p'][1]
        assert len(a) == len(b) == n_utts
        frames = 0
        for i in range(n_utts):
            assert a[i] is not None and b[i] is not None and a[i].shape == b[i].shape, i
            assert torch.equal(a[i].to(dev), b[i].to(dev)), f'utterance {i} differs'
            frames += a[i].shape[1]
        print(f'p2p gather == nccl gather for {n_utts} utterances, {frames} frames, world {world}: '
              f'nccl {res["nccl"][0] * 1e3:.1f} ms, peer-window {res["p2p"][0] * 1e3:.1f} ms '
              f'({frames / res["p2p"][0] / 1e6:.2f} M frames/s end to end incl. host bucketing)', flush=True)
    window.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
