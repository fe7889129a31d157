"""Mirror of the reference's ``models/common_layers.py`` surface.

The classes keep the reference's constructor arguments and parameter names so
``state_dict()`` / ``load_state_dict()`` are interchangeable with upstream
checkpoints (models/common_layers.py:22-84), but they hold parameters only: all
arithmetic runs in the sm_100a extension through the C ABI.  ``LengthRegulator``
is a complete operator on its own (models/common_layers.py:7-19); ``CBHG`` and
the predictors execute through their owning model's handle
(``ForwardTacotron.run_cbhg`` / ``run_series_predictor``).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.nn as nn

from .. import _lib


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f'{what}: tensor is on {t.device}; forwardtacotron_b200 runs on sm_100a GPUs only '
                           '(there is no CPU fallback)')


class LengthRegulator(nn.Module):
    """Expands phoneme rows by rounded durations; bit-exact against the reference.

    ``forward`` clamps ``dur`` in place like the reference does (line 13)."""

    def forward(self, x: torch.Tensor, dur: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, 'LengthRegulator')
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError('LengthRegulator supports float32 and bfloat16 payloads')
        if dur.dtype != torch.float32 or not dur.is_contiguous():
            raise TypeError('dur must be a contiguous float32 tensor (it is clamped in place)')
        B, T, Cn = x.shape
        cum, total = self.plan(dur)
        L = int(total.max().item())  # the one device->host read: output size is data dependent
        return self.expand(x, cum, L)

    @staticmethod
    def plan(dur: torch.Tensor):
        lib = _lib.lib()
        B, T = dur.shape
        cum = torch.empty((B, T), dtype=torch.int32, device=dur.device)
        total = torch.empty((B,), dtype=torch.int32, device=dur.device)
        with torch.cuda.device(dur.device):
            _lib.check(lib.ftb_length_plan(_lib.ptr(dur), _lib.ptr(cum), _lib.ptr(total), B, T,
                                           _lib.current_stream(dur.device)))
        return cum, total

    @staticmethod
    def expand(x: torch.Tensor, cum: torch.Tensor, L: int) -> torch.Tensor:
        lib = _lib.lib()
        B, T, Cn = x.shape
        x = x.contiguous()
        out = torch.empty((B, L, Cn), dtype=x.dtype, device=x.device)
        if L == 0:
            return out
        with torch.cuda.device(x.device):
            _lib.check(lib.ftb_length_expand(_lib.ptr(x), _lib.ptr(cum), _lib.ptr(out), B, T, L, Cn,
                                             x.element_size(), _lib.current_stream(x.device)))
        return out


class HighwayNetwork(nn.Module):
    def __init__(self, size: int) -> None:
        super().__init__()
        self.W1 = nn.Linear(size, size)
        self.W2 = nn.Linear(size, size)
        with torch.no_grad():
            self.W1.bias.zero_()  # models/common_layers.py:28


class BatchNormConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel: int, relu: bool = True) -> None:
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel, stride=1, padding=kernel // 2, bias=False)
        self.bnorm = nn.BatchNorm1d(out_channels)
        self.relu = relu


class CBHG(nn.Module):
    def __init__(self, K: int, in_channels: int, channels: int, proj_channels: List[int], num_highways: int,
                 dropout: float = 0.5) -> None:
        super().__init__()
        self.dropout = dropout  # identity at inference
        self.bank_kernels = list(range(1, K + 1))
        self.conv1d_bank = nn.ModuleList(BatchNormConv(in_channels, channels, k) for k in self.bank_kernels)
        self.conv_project1 = BatchNormConv(K * channels, proj_channels[0], 3)
        self.conv_project2 = BatchNormConv(proj_channels[0], proj_channels[1], 3, relu=False)
        self.pre_highway = nn.Linear(proj_channels[-1], channels, bias=False)
        self.highways = nn.ModuleList(HighwayNetwork(channels) for _ in range(num_highways))
        self.rnn = nn.GRU(channels, channels, batch_first=True, bidirectional=True)


class NativeModel(nn.Module):
    """Shared plumbing of the two TTS models: native handle lifetime + workspace.

    Native state is kept per (device, CUDA stream) "lane": a C handle (packed weight copy, internal side streams and
    events) is not re-entrant, so calling ``generate`` under several ``torch.cuda.stream(...)`` contexts gives every
    stream its own handle and workspace and the calls overlap on the GPU -- the sequential recurrences of one batch
    leave most SMs idle for another batch's GEMMs (measured +23 % / +30 % throughput with 2 / 3 streams).
    Parameters are shared; each lane costs one packed weight copy (~50 MB) plus its workspace."""

    _create_fn = ''
    _destroy_fn = ''

    def __init__(self) -> None:
        super().__init__()
        self._lanes = {}       # (device index, stream handle) -> {'handle', 'keep', 'ws'}
        self.gemm_mode = 0     # see ftb_ft_config / ftb_fp_config in include/ftb200.h

    # -- invalidation: any re-materialisation of the parameters drops the packed copies
    def _apply(self, fn, *args, **kwargs):
        self._drop_handle()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._drop_handle()
        return super().load_state_dict(*args, **kwargs)

    def refresh(self) -> None:
        """Re-pack the weights on the next call (use after in-place parameter edits)."""
        self._drop_handle()

    def _drop_handle(self) -> None:
        lanes = self.__dict__.get('_lanes') or {}
        if lanes:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            for lane in lanes.values():
                getattr(_lib.lib(), self._destroy_fn)(lane['handle'])
        self.__dict__['_lanes'] = {}

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _config_struct(self):
        raise NotImplementedError

    def _lane_key(self, device: torch.device):
        return (device.index or 0, torch.cuda.current_stream(device).cuda_stream)

    @property
    def _handle(self):
        """Handle of the lane of the current stream on the parameters' device (None before the first call)."""
        try:
            device = next(self.parameters()).device
        except StopIteration:
            return None
        if device.type != 'cuda':
            return None
        lane = self._lanes.get(self._lane_key(device))
        return lane['handle'] if lane else None

    def _get_handle(self, device: torch.device):
        key = self._lane_key(device)
        lane = self._lanes.get(key)
        if lane is not None:
            return lane['handle']
        lib = _lib.lib()
        sd = self.state_dict()
        for k, v in sd.items():
            _require_cuda(v, f'parameter {k}')
        table, keep = _lib.tensor_table(sd.items())
        cfg = self._config_struct()
        out = C.c_void_p()
        torch.cuda.synchronize(device)
        with torch.cuda.device(device):
            _lib.check(getattr(lib, self._create_fn)(C.byref(cfg), table, len(table), device.index or 0,
                                                     C.byref(out)))
        self._lanes[key] = {'handle': out, 'keep': (table, keep, cfg), 'ws': None}
        if len(self._lanes) >= 2 and hasattr(lib, 'ftb_ft_set_option') and self._create_fn == 'ftb_ft_create':
            # a second stream means batches in flight: trade a little single-call latency of the decoder LSTM for
            # fewer SMs held during its recurrence -- an option of THIS model's handles, not a process-wide setting
            # (include/ftb200.h, FTB_OPT_LSTM_MIN_CHUNK)
            for lane in self._lanes.values():
                _lib.check(lib.ftb_ft_set_option(lane['handle'], _lib.FTB_OPT_LSTM_MIN_CHUNK, 32))
        return out

    def lane_streams(self, device, n: int):
        """``n`` long-lived CUDA streams for batches in flight.  Lanes are keyed by stream, so callers that overlap
        batches must reuse these instead of creating streams per call (each new stream would cost a packed weight
        copy and a workspace)."""
        device = torch.device(device)
        pool = self.__dict__.setdefault('_stream_pool', {})
        have = pool.setdefault(device.index or 0, [])
        while len(have) < n:
            have.append(torch.cuda.Stream(device))
        return have[:n]

    def _get_workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        lane = self._lanes[self._lane_key(device)]
        ws = lane['ws']
        if ws is None or ws.device != device or ws.numel() < nbytes:
            lane['ws'] = None
            ws = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=device)
            lane['ws'] = ws
        return ws

    def forward(self, *args, **kwargs):
        raise NotImplementedError('teacher-forced training forward() is outside the hot path this package '
                                  'implements (SURVEY 8f-4); use generate()')
