"""Peer-mapped result window: the fused form of the path's only exchange step.

``generate`` ends with a GEMM (``post_proj``) whose epilogue writes ``mel_post`` (B, 80, L).  In a sharded run the
plain design gathers those tensors afterwards with NCCL (utils/batching.gather_mels).  Here rank ``dst`` allocates one
window in its HBM, every rank of the node maps it through CUDA IPC (ftb_ipc_*), and each rank's epilogue stores its
result tile straight into its slot over NVLink / NVSwitch: compute and "collective" are one kernel, there is no gather
pass and no staging copy.  The only collective left is a tiny all-gather of the slot sizes (3 integers per rank and
step).

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

    def __init__(self, ptr: int, shape: Tuple[int, int, int]):
        self._ptr, self.shape = ptr, tuple(shape)

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
        self.cursor = 0
        self.slots: List[Tuple[int, int, Tuple[int, int, int]]] = []  # on dst: (rank, byte offset, shape)

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
        self.cursor = 0
        self.slots = []

    def alloc(self, B: int, n_mels: int, L: int):
        """COLLECTIVE: every rank calls it once per step (ranks without work pass B = 0).  Returns this rank's slot:
        a float32 (B, n_mels, L) tensor view on the owner, a ``PeerSlot`` elsewhere, None for an empty slot."""
        mine = torch.tensor([B, n_mels, L], dtype=torch.long, device=self.device)
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine, group=self.group)
        sizes = [tuple(int(v) for v in t.tolist()) for t in allv]
        my_off = None
        off = self.cursor
        for r, (b, m, l) in enumerate(sizes):
            nbytes = (b * m * l * 4 + 255) // 256 * 256
            if r == self.rank:
                my_off = off
            if self.rank == self.dst and b > 0:
                self.slots.append((r, off, (b, m, l)))
            off += nbytes
        if off > self.capacity:
            raise RuntimeError(f'PeerWindow overflow: {off} > {self.capacity} bytes')
        self.cursor = off
        if B == 0:
            return None
        if self.rank == self.dst:
            return self._view[my_off:my_off + B * n_mels * L * 4].view(torch.float32).view(B, n_mels, L)
        return PeerSlot(self._ptr.value + my_off, (B, n_mels, L))

    def collect(self) -> Optional[List[Tuple[int, torch.Tensor]]]:
        """COLLECTIVE: waits until every rank's kernels have finished, then returns on ``dst`` the list of
        (producing rank, tensor view) in allocation order (views into the window, no copy); None elsewhere."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        if self.rank != self.dst:
            return None
        out = []
        for r, off, shape in self.slots:
            n = shape[0] * shape[1] * shape[2]
            out.append((r, self._view[off:off + n * 4].view(torch.float32).view(*shape)))
        return out
