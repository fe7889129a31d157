"""Host-side batching of the gen_forward loop (row a14) and the multi-process path (row e): bucketing, rank
sharding and the final variable-length gather, the latter with world_size 2 on the gloo backend."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from forwardtacotron_b200.utils import batching


def make_utts(n, seed=0, lo=3, hi=60):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi, (n,), generator=g).tolist()
    return [torch.randint(1, 135, (L,), generator=g).tolist() for L in lens]


def test_buckets_cover_every_utterance_once_and_respect_limits():
    utts = make_utts(101)
    batches = batching.bucket_by_length(utts, max_tokens=400, max_batch=16)
    seen = sorted(int(i) for b in batches for i in b.index)
    assert seen == list(range(101))
    for b in batches:
        B, T = b.tokens.shape
        assert B <= 16 and (B * T <= 400 or B == 1)
        assert T == int(b.lengths.max())
        for r, i in enumerate(b.index.tolist()):
            L = len(utts[i])
            assert b.tokens[r, :L].tolist() == utts[i] and torch.all(b.tokens[r, L:] == batching.PAD_ID)
    # sorted by length: padding overhead stays small
    pad = sum(b.tokens.numel() for b in batches) / sum(len(u) for u in utts)
    assert pad < 1.15


def test_empty_utterance_is_an_error():
    with pytest.raises(ValueError):
        batching.bucket_by_length([[1, 2], []])


@pytest.mark.parametrize('world', [1, 2, 4, 8])
def test_sharding_is_a_balanced_partition(world):
    batches = batching.bucket_by_length(make_utts(300, seed=3), max_tokens=512)
    parts = [batching.shard_for_rank(batches, r, world) for r in range(world)]
    ids = sorted(int(i) for p in parts for b in p for i in b.index)
    assert ids == list(range(300))
    loads = [sum(b.tokens.numel() for b in p) for p in parts]
    assert max(loads) - min(loads) <= max(b.tokens.numel() for b in batches)


def test_gather_without_process_group():
    mels = [torch.randn(80, 5), torch.randn(80, 9)]
    out = batching.gather_mels(mels, [2, 0], 3)
    assert out[1] is None and torch.equal(out[0], mels[1]) and torch.equal(out[2], mels[0])


class FakeModel(torch.nn.Module):
    """Stands in for the CUDA model with the reference's NO-MASK semantics: every token -- pad tokens included -- lasts
    2 frames (+1 for the first), so a padded row is longer than its real tokens' frames and must be cut."""
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def generate(self, x, alpha=1.0, **kw):
        B, T = x.shape
        dur = torch.full((B, T), 2.0)
        dur[:, 0] = 3.0
        L = 2 * T + 1
        mel = torch.zeros(B, 80, L)
        for b in range(B):
            mel[b] = x[b].float().sum() + torch.arange(L)[None, :]
        return {'mel_post': mel, 'mel': mel, 'dur': dur, 'mel_len': torch.full((B,), L)}


class FakeRaggedModel(FakeModel):
    """Offers generate_ragged: rows are computed from their own tokens only (mel_len per row)."""
    def generate_ragged(self, x, lengths, alpha=1.0, **kw):
        out = self.generate(x, alpha)
        out['mel_len'] = torch.as_tensor(lengths) * 2 + 1
        out['mel_post'] = out['mel_post'] + 1000.0          # marks the path taken
        return out


def test_corpus_rows_are_cut_at_the_frames_of_their_real_tokens():
    utts = make_utts(23, seed=8)
    for model, marker in ((FakeModel(), 0.0), (FakeRaggedModel(), 1000.0)):
        out = batching.synthesize_corpus(model, utts, max_tokens=200, device=torch.device('cpu'))
        for i, u in enumerate(utts):
            assert out[i].shape == (80, 2 * len(u) + 1)
            assert float(out[i][0, 0]) == float(sum(u)) + marker
    with pytest.raises(ValueError):
        batching.synthesize_corpus(FakeModel(), utts, device=torch.device('cpu'), exact=True)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        utts = make_utts(37, seed=5)
        out = batching.synthesize_corpus(FakeModel(), utts, max_tokens=200, device=torch.device('cpu'))
        if rank == 0:
            ok = out is not None and len(out) == 37
            for i, u in enumerate(utts):
                want_len = 2 * len(u) + 1
                ok = ok and out[i] is not None and out[i].shape == (80, want_len) and float(out[i][0, 0]) == float(sum(u))
            q.put(bool(ok))
        else:
            q.put(out is None)
    finally:
        dist.destroy_process_group()


def test_sharded_synthesis_and_gather_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert res == [True, True] and all(p.exitcode == 0 for p in procs)
