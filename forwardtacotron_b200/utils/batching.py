"""Batched replacement for the reference's per-sentence loop (gen_forward.py:106-118, B = 1 there).

Host-side logic only (no kernels): length-bucketed padded batches, utterance sharding over
``torch.distributed`` ranks, and the one exchange step of the path — the final gather of the
variable-length mels.  Utterances are independent, so there is no collective inside ``generate``.

Two batch semantics (``synthesize_corpus(exact=...)``):

* ``exact=True`` (default for models that offer ``generate_ragged``, i.e. ForwardTacotron): every row carries its own
  token count into the kernels — convs see zero padding beyond the row's end, recurrences run over the row's own
  length, the duration fallback is decided per row — so each utterance gets exactly the mel the reference's
  one-sentence-per-call loop produces, at batched throughput.
* ``exact=False``: the reference's no-mask semantics on the padded batch (pad id 0 = ``'_'`` is an ordinary symbol for
  ForwardTacotron: pad tokens receive durations and would expand into frames).  Rows are cut after the frames
  of their REAL tokens, so pad frames never reach the output, but the real frames still feel the padding through the
  bidirectional recurrences.  Kept for FastPitch (whose prenet masks pad keys itself) and for measurements.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

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


def _dist_state(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(group), dist.get_world_size(group)
    return None, 0, 1


def gather_mels(mels: Sequence[torch.Tensor], index: Sequence[int], n_total: int, dst: int = 0,
                group=None) -> Optional[List[Optional[torch.Tensor]]]:
    """The path's single exchange step: collect the variable-length ``(n_mels, L_i)`` results of every rank on ``dst``
    in the caller's original order.  ``mels[k]`` belongs to utterance ``index[k]``.

    NCCL has no gatherv, so: the per-utterance (index, n_mels, L) metadata travels as one small object
    all-gather, then every rank packs its mels into ONE flat buffer and sends exactly those bytes to ``dst`` — grouped
    point-to-point (``batch_isend_irecv`` = ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd).  Nothing is padded to
    a global maximum, no rank but ``dst`` receives anything, and ``dst`` hands out views of the received buffers.
    Returns the list on ``dst`` (``None`` for utterances nobody produced) and ``None`` elsewhere.  Works without an
    initialised process group (single process) and with the gloo backend (CPU tests)."""
    if len(mels) != len(index):
        raise ValueError('mels and index must have the same length')
    dist, rank, world = _dist_state(group)
    if dist is None:
        out: List[Optional[torch.Tensor]] = [None] * n_total
        for m, i in zip(mels, index):
            out[int(i)] = m
        return out
    if dist.get_backend(group) == 'nccl':
        dev = mels[0].device if len(mels) else torch.device('cuda', torch.cuda.current_device())
    else:
        dev = torch.device('cpu')
    info = [(int(i), int(m.shape[0]), int(m.shape[1])) for m, i in zip(mels, index)]  # host-side, no device reads
    infos: List[Optional[list]] = [None] * world
    dist.all_gather_object(infos, info, group=group)
    sizes = [sum(c * l for _, c, l in inf) for inf in infos]
    flat = torch.cat([m.reshape(-1) for m in mels]).to(device=dev, dtype=torch.float32) if len(mels) \
        else torch.empty(0, dtype=torch.float32, device=dev)
    ops, bufs = [], {}
    if rank == dst:
        bufs[rank] = flat
        for r in range(world):
            if r != dst and sizes[r] > 0:
                bufs[r] = torch.empty(sizes[r], dtype=torch.float32, device=dev)
                ops.append(dist.P2POp(dist.irecv, bufs[r], _global_rank(dist, r, group), group))
    elif sizes[rank] > 0:
        ops.append(dist.P2POp(dist.isend, flat, _global_rank(dist, dst, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    out = [None] * n_total
    for r, inf in enumerate(infos):
        if not inf:
            continue
        parts = bufs[r].split([c * l for _, c, l in inf])
        for (i, c, l), p in zip(inf, parts):
            out[i] = p.view(c, l)
    return out


def _global_rank(dist, r: int, group) -> int:
    return r if group is None else dist.get_global_rank(group, r)


def _row_frames(out, batch: Batch, exact: bool) -> List[int]:
    """Frames each row's REAL tokens produced (one device -> host read per batch)."""
    if exact:
        return out['mel_len'].tolist()
    # no-mask batch: pad tokens expanded into frames at the END of the row; count only what the real tokens produced
    r = (out['dur'].clamp(min=0) + 0.5).long()
    pos = torch.arange(r.shape[1], device=r.device)[None, :] < batch.lengths.to(r.device)[:, None]
    return (r * pos).sum(1).tolist()


def synthesize_corpus(model, utterances: Sequence[Sequence[int]], alpha: float = 1.0, max_tokens: int = 16384,
                      max_batch: int = 256, key: str = 'mel_post', device=None, window=None, in_flight: int = 3,
                      exact: Optional[bool] = None, gather: bool = True,
                      **callbacks) -> Optional[List[Optional[torch.Tensor]]]:
    """gen_forward.py's loop, batched and sharded: bucket, run the model on this rank's batches, cut every row at its own
    frame count and collect on rank 0.  ``model`` is a ForwardTacotron / FastPitch mirror already on its device.
    ``exact`` (default: True when the model offers ``generate_ragged``): every utterance gets the mel of its own
    one-sentence ``generate`` call (module docstring).  ``in_flight`` batches are issued round-robin on as many CUDA
    streams (the model keeps one native lane per stream), so one batch's GEMMs fill the SMs its neighbour's recurrences
    leave idle.

    ``window`` (utils/peer_window.PeerWindow): instead of the NCCL gather, every rank's last GEMM stores ``mel_post``
    directly into rank 0's memory over NVLink; rank 0 slices the results out of the window.
    ``gather=False`` skips the exchange and returns this rank's own results (others ``None``): the baseline the cost of
    the exchange is measured against."""
    _, rank, world = _dist_state()
    device = device or next(model.parameters()).device
    if exact is None:
        exact = hasattr(model, 'generate_ragged')
    if exact and not hasattr(model, 'generate_ragged'):
        raise ValueError(f'{type(model).__name__} has no generate_ragged(); use exact=False (or max_batch=1)')
    mine = shard_for_rank(bucket_by_length(utterances, max_tokens, max_batch), rank, world)
    on_gpu = torch.device(device).type == 'cuda'
    n_streams = max(1, min(int(in_flight), len(mine))) if on_gpu else 1
    if n_streams > 1:  # the model's own long-lived streams: one native lane (packed weights, workspace) per stream
        streams = (model.lane_streams(device, n_streams) if hasattr(model, 'lane_streams')
                   else [torch.cuda.Stream(device) for _ in range(n_streams)])
        here = torch.cuda.current_stream(device)
        for s in streams:
            s.wait_stream(here)
    else:
        streams = [None]
    if window is not None:
        window.reset()
    alloc = window.alloc if window is not None else None

    def run(b: Batch):
        tok = b.tokens.to(device, non_blocking=True)
        kw = dict(callbacks)
        if alloc is not None:
            kw['mel_post_alloc'] = alloc
        if exact:
            return model.generate_ragged(tok, b.lengths, alpha, **kw)
        return model.generate(tok, alpha, **kw)

    outs = []
    for k, b in enumerate(mine):
        s = streams[k % n_streams]
        if s is None:
            outs.append((b, run(b)))
        else:
            with torch.cuda.stream(s):
                outs.append((b, run(b)))
    if n_streams > 1:
        for s in streams:
            torch.cuda.current_stream(device).wait_stream(s)
    mels: List[torch.Tensor] = []
    index: List[int] = []
    for b, out in outs:
        if n_streams > 1 and isinstance(out[key], torch.Tensor):
            out[key].record_stream(torch.cuda.current_stream(device))  # produced on a side stream, consumed here
        rows, lens = b.index.tolist(), _row_frames(out, b, exact)
        if window is not None:
            window.note_rows(out[key], rows, lens)
        else:
            for r, (i, n) in enumerate(zip(rows, lens)):
                mels.append(out[key][r, :, :n])
                index.append(i)
    if window is not None:
        return window.collect(len(utterances))
    if not gather:
        local: List[Optional[torch.Tensor]] = [None] * len(utterances)
        for m, i in zip(mels, index):
            local[i] = m
        return local
    return gather_mels(mels, index, len(utterances))
