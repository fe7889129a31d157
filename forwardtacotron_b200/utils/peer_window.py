"""Peer-mapped result window: the fused form of the path's only exchange step.

``generate`` ends with a GEMM (``post_proj``) whose epilogue writes ``mel_post`` (B, 80, L).  In a sharded run the
plain design gathers those tensors afterwards with NCCL (utils/batching.gather_mels).  Here rank ``dst`` allocates one
window in its HBM, every rank of the node maps it through CUDA IPC (ftb_ipc_*), and each rank's epilogue stores its
result tile straight into the window over NVLink / NVSwitch: compute and "collective" are one kernel, there is no gather
pass and no staging copy.

The window is cut into one REGION per rank and every rank bump-allocates inside its own region, so ``alloc`` needs no
communication at all: it is safe to call from inside ``generate`` with several batches in flight, and a rank that raises
cannot leave its peers waiting in a collective.  The only collective is ``collect`` at the end of the corpus: one small
object all-gather of the slot metadata (offsets, shapes, utterance ids) after every rank has synchronised its device.

Single node only (CUDA IPC); one process per GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import _lib


class PeerSlot:
    """What a non-owning rank gets back from ``alloc``: a (B, n_mels, L) float32 region of the owner's HBM that this
    rank's kernels can store to.  It quacks enough like a tensor for ``synthesize`` to take it as the output."""
    is_cuda, dtype = True, torch.float32

    def __init__(self, ptr: int, shape: Tuple[int, int, int], offset: int):
        self._ptr, self.shape, self.offset = ptr, tuple(shape), offset

    def data_ptr(self) -> int:
        return self._ptr

    def is_contiguous(self) -> bool:
        return True


class _Raw:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


class PeerWindow:
    def __init__(self, capacity_bytes: int, dst: int = 0, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('PeerWindow needs an initialised torch.distributed process group')
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(capacity_bytes)
        self.region = self.capacity // self.world // 256 * 256      # bytes each rank may fill
        self.device = torch.device('cuda', torch.cuda.current_device())
        lib = _lib.lib()
        self._ptr = C.c_void_p()
        payload = [None]
        if self.rank == dst:
            handle = (C.c_ubyte * 64)()
            _lib.check(lib.ftb_ipc_alloc(self.capacity, self.device.index, C.byref(self._ptr), handle))
            payload = [bytes(handle)]
            # the owner reads results through an ordinary tensor view of the window
            self._view = torch.as_tensor(_Raw(self._ptr.value, self.capacity), device=self.device)
        dist.broadcast_object_list(payload, src=dst, group=group)
        if self.rank != dst:
            handle = (C.c_ubyte * 64).from_buffer_copy(payload[0])
            _lib.check(lib.ftb_ipc_open(handle, self.device.index, C.byref(self._ptr)))
            self._view = None
        self.reset()

    def close(self) -> None:
        if self._ptr:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody un-maps / frees while a peer may still store
            if self.rank != self.dst:
                _lib.check(_lib.lib().ftb_ipc_release(self._ptr, 0))
            dist.barrier(group=self.group)
            if self.rank == self.dst:
                self._view = None
                _lib.check(_lib.lib().ftb_ipc_release(self._ptr, 1))
            self._ptr = C.c_void_p()

    def reset(self) -> None:
        """Start a new corpus: the regions are reused from their beginning.  Results handed out by an earlier
        ``collect`` are copies, so they stay valid."""
        self.cursor = self.rank * self.region
        self.meta: List[Tuple[int, Tuple[int, int, int], List[int], List[int]]] = []

    def alloc(self, B: int, n_mels: int, L: int):
        """LOCAL (no communication): reserves a float32 (B, n_mels, L) slot in this rank's region of the window.
        Returns a tensor view on the owner, a ``PeerSlot`` elsewhere."""
        nbytes = (B * n_mels * L * 4 + 255) // 256 * 256
        off = self.cursor
        if off + nbytes > (self.rank + 1) * self.region:
            raise RuntimeError(f'PeerWindow: region of rank {self.rank} is full ({self.region} bytes per rank); '
                               'allocate a larger window')
        self.cursor = off + nbytes
        if self.rank == self.dst:
            t = self._view[off:off + B * n_mels * L * 4].view(torch.float32).view(B, n_mels, L)
            t._ftb_window_offset = off
            return t
        return PeerSlot(self._ptr.value + off, (B, n_mels, L), off)

    def note_rows(self, slot, rows: List[int], frames: List[int]) -> None:
        """Records which utterances (and how many valid frames each) the slot's rows hold."""
        off = slot.offset if isinstance(slot, PeerSlot) else slot._ftb_window_offset
        self.meta.append((int(off), tuple(int(v) for v in slot.shape), [int(r) for r in rows], [int(f) for f in frames]))

    def collect(self, n_total: int) -> Optional[List[Optional[torch.Tensor]]]:
        """COLLECTIVE (the only one): every rank waits for its own kernels, the slot metadata is all-gathered, and
        ``dst`` returns the per-utterance ``(n_mels, L_i)`` results in the caller's order (copies out of the window, so
        they survive the next ``reset``); None elsewhere."""
        torch.cuda.synchronize(self.device)         # this rank's peer stores have landed in the owner's HBM
        metas: List[Optional[list]] = [None] * self.world
        dist.all_gather_object(metas, self.meta, group=self.group)
        if self.rank != self.dst:
            return None
        out: List[Optional[torch.Tensor]] = [None] * n_total
        for r, meta in enumerate(metas):
            if not meta:
                continue
            # ONE copy per producing rank (the used part of its region), then views: the results survive reset()
            lo = r * self.region
            hi = max(off + shape[0] * shape[1] * shape[2] * 4 for off, shape, _, _ in meta)
            keep = self._view[lo:hi].clone()
            for off, shape, rows, frames in meta:
                n = shape[0] * shape[1] * shape[2]
                t = keep[off - lo:off - lo + n * 4].view(torch.float32).view(*shape)
                for row, (i, f) in enumerate(zip(rows, frames)):
                    out[i] = t[row, :, :f]
        return out
