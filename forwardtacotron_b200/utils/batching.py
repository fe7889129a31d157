"""Batched replacement for the reference's per-sentence loop (gen_forward.py:106-118, B = 1 there).

Host-side logic only (no kernels): length-bucketed padded batches, utterance sharding over
``torch.distributed`` ranks, and the one exchange step of the path — the final gather of the
variable-length mels.  Utterances are independent, so there is no collective inside ``generate``.

Reference semantics that are kept on purpose: the pad id is 0 = ``'_'`` (utils/text/symbols.py),
and pad tokens are ordinary symbols for ForwardTacotron — they receive durations and expand into frames.
Bucketing by length keeps the padding (and that effect) small; it does not remove it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

PAD_ID = 0


@dataclass
class Batch:
    """One padded batch: ``tokens`` (B, T) int64, ``lengths`` (B,) token counts, ``index`` (B,) positions
    of the rows in the caller's original utterance list."""
    tokens: torch.Tensor
    lengths: torch.Tensor
    index: torch.Tensor


def bucket_by_length(utterances: Sequence[Sequence[int]], max_tokens: int = 16384, max_batch: int = 256) -> List[Batch]:
    """Sort by token count and cut into batches whose padded size B*T stays <= ``max_tokens`` (and B <=
    ``max_batch``).  Deterministic; every utterance appears in exactly one batch."""
    if max_tokens <= 0 or max_batch <= 0:
        raise ValueError('max_tokens and max_batch must be positive')
    order = sorted(range(len(utterances)), key=lambda i: (len(utterances[i]), i))
    batches: List[Batch] = []
    cur: List[int] = []
    for i in order:
        n = len(utterances[i])
        if n == 0:
            raise ValueError(f'utterance {i} is empty')
        # sorted ascending: the candidate is the longest row of the batch it would join
        if cur and (len(cur) + 1 > max_batch or (len(cur) + 1) * n > max_tokens):
            batches.append(_pad(utterances, cur))
            cur = []
        cur.append(i)
    if cur:
        batches.append(_pad(utterances, cur))
    return batches


def _pad(utterances, idx: List[int]) -> Batch:
    T = max(len(utterances[i]) for i in idx)
    tok = torch.full((len(idx), T), PAD_ID, dtype=torch.long)
    for r, i in enumerate(idx):
        tok[r, :len(utterances[i])] = torch.as_tensor(list(utterances[i]), dtype=torch.long)
    return Batch(tok, torch.tensor([len(utterances[i]) for i in idx], dtype=torch.long), torch.tensor(idx, dtype=torch.long))


def shard_for_rank(batches: Sequence[Batch], rank: int, world_size: int) -> List[Batch]:
    """Deal the batches over the ranks so the padded token counts (~ frames, ~ time) balance: heaviest batch
    first, each to the currently lightest rank (LPT).  Every rank computes the same assignment."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    load = [0] * world_size
    mine: List[Batch] = []
    for j in sorted(range(len(batches)), key=lambda j: (-batches[j].tokens.numel(), j)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += batches[j].tokens.numel()
        if r == rank:
            mine.append(batches[j])
    return mine


def gather_mels(mels: Sequence[torch.Tensor], index: Sequence[int], n_total: int, dst: int = 0,
                group=None) -> Optional[List[Optional[torch.Tensor]]]:
    """The path's single collective: collect variable-length ``(n_mels, L_i)`` results of every rank on
    ``dst`` in the caller's original order.  ``mels[k]`` belongs to utterance ``index[k]``.

    NCCL / gloo have no gatherv, so: all-gather the per-rank (count, max L), then one all-gather of the
    rank's mels padded to the global max (payload is tiny next to the compute: 80 x L floats per utterance).
    Returns the list on ``dst`` (``None`` for utterances nobody produced) and ``None`` elsewhere.
    Works without an initialised process group (single process)."""
    import torch.distributed as dist
    if len(mels) != len(index):
        raise ValueError('mels and index must have the same length')
    if not (dist.is_available() and dist.is_initialized()):
        out: List[Optional[torch.Tensor]] = [None] * n_total
        for m, i in zip(mels, index):
            out[int(i)] = m
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = mels[0].device if len(mels) else torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' \
        else torch.device('cpu')
    n_mels = mels[0].shape[0] if len(mels) else 0
    meta = torch.tensor([len(mels), max((m.shape[1] for m in mels), default=0), n_mels], dtype=torch.long, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    cnt = max(int(m[0]) for m in metas)
    lmax = max(int(m[1]) for m in metas)
    n_mels = max(int(m[2]) for m in metas)
    if cnt == 0:
        return [None] * n_total if rank == dst else None
    payload = torch.zeros((cnt, n_mels, lmax), dtype=torch.float32, device=dev)
    info = torch.full((cnt, 2), -1, dtype=torch.long, device=dev)  # (utterance index, L)
    for k, (m, i) in enumerate(zip(mels, index)):
        payload[k, :, :m.shape[1]] = m
        info[k, 0], info[k, 1] = int(i), m.shape[1]
    payloads = [torch.zeros_like(payload) for _ in range(world)]
    infos = [torch.zeros_like(info) for _ in range(world)]
    dist.all_gather(payloads, payload, group=group)
    dist.all_gather(infos, info, group=group)
    if rank != dst:
        return None
    out = [None] * n_total
    for p, inf in zip(payloads, infos):
        for k in range(cnt):
            i, L = int(inf[k, 0]), int(inf[k, 1])
            if i >= 0:
                out[i] = p[k, :, :L].clone()
    return out


def synthesize_corpus(model, utterances: Sequence[Sequence[int]], alpha: float = 1.0, max_tokens: int = 16384,
                      max_batch: int = 256, key: str = 'mel_post', device=None, window=None, in_flight: int = 3,
                      **callbacks) -> Optional[List[Optional[torch.Tensor]]]:
    """gen_forward.py's loop, batched and sharded: bucket, run ``model.generate`` on this rank's batches, cut every
    row at its own frame count (``mel_len``) and gather on rank 0.  ``model`` is a ForwardTacotron / FastPitch
    mirror already on its device.  ``in_flight`` batches are issued round-robin on as many CUDA streams (the model
    keeps one native lane per stream), so one batch's GEMMs fill the SMs its neighbour's recurrences leave idle.

    ``window`` (utils/peer_window.PeerWindow): instead of the NCCL gather, every rank's last GEMM stores ``mel_post``
    directly into rank 0's memory over NVLink; rank 0 slices views out of the window."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    device = device or next(model.parameters()).device
    mels: List[torch.Tensor] = []
    index: List[int] = []
    if window is not None:
        return _synthesize_into_window(model, utterances, alpha, max_tokens, max_batch, device, window, rank, world,
                                       callbacks)
    mine = shard_for_rank(bucket_by_length(utterances, max_tokens, max_batch), rank, world)
    on_gpu = torch.device(device).type == 'cuda'
    n_streams = max(1, min(int(in_flight), len(mine))) if on_gpu else 1
    if n_streams > 1:  # the model's own long-lived streams: one native lane (packed weights, workspace) per stream
        streams = (model.lane_streams(device, n_streams) if hasattr(model, 'lane_streams')
                   else [torch.cuda.Stream(device) for _ in range(n_streams)])
    else:
        streams = [None]
    if n_streams > 1:
        here = torch.cuda.current_stream(device)
        for s in streams:
            s.wait_stream(here)
    outs = []
    for k, b in enumerate(mine):
        s = streams[k % n_streams]
        if s is None:
            outs.append((b, model.generate(b.tokens.to(device), alpha, **callbacks)))
        else:
            with torch.cuda.stream(s):
                outs.append((b, model.generate(b.tokens.to(device, non_blocking=True), alpha, **callbacks)))
    if n_streams > 1:
        for s in streams:
            torch.cuda.current_stream(device).wait_stream(s)
    for b, out in outs:
        if n_streams > 1:
            out[key].record_stream(torch.cuda.current_stream(device))  # produced on a side stream, consumed here
        lens = out['mel_len'].tolist()
        for r, i in enumerate(b.index.tolist()):
            mels.append(out[key][r, :, :lens[r]])
            index.append(i)
    return gather_mels(mels, index, len(utterances))


def _synthesize_into_window(model, utterances, alpha, max_tokens, max_batch, device, window, rank, world, callbacks):
    import torch.distributed as dist
    batches = bucket_by_length(utterances, max_tokens, max_batch)
    mine = shard_for_rank(batches, rank, world)
    steps = torch.tensor([len(mine)], dtype=torch.long, device=device)
    dist.all_reduce(steps, op=dist.ReduceOp.MAX)
    window.reset()
    meta = []  # per produced slot on this rank: (utterance indices, frame counts)
    for k in range(int(steps)):
        if k < len(mine):
            out = model.generate(mine[k].tokens.to(device), alpha, mel_post_alloc=window.alloc, **callbacks)
            meta.append((mine[k].index.tolist(), out['mel_len'].tolist()))
        else:
            window.alloc(0, 0, 0)  # keep the collective call sequence aligned
    metas = [None] * world
    dist.all_gather_object(metas, meta)
    slots = window.collect()
    if slots is None:
        return None
    out = [None] * len(utterances)
    seen = [0] * world
    for r, t in slots:
        idx, lens = metas[r][seen[r]]
        seen[r] += 1
        for row, (i, L) in enumerate(zip(idx, lens)):
            out[i] = t[row, :, :L]
    return out
