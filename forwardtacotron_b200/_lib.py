"""ctypes binding of include/ftb200.h (the C ABI of csrc/libftb200.so).

There is no CPU fallback anywhere in this package: if the shared library is
missing it is built in-tree with nvcc, and if that is impossible importing
``lib()`` raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Iterable, List, Tuple

_LIB = None
_LIB_PATH = Path(__file__).resolve().parent / 'csrc' / 'libftb200.so'

FTB_F32, FTB_I64, FTB_BF16, FTB_I32 = 0, 1, 2, 3
FTB_OPT_OVERLAP_PRENET = 1
FTB_OPT_SERIALIZE = 2
FTB_OPT_DUR_SIMT = 3
FTB_OPT_UNFUSED_TAIL = 4
FTB_OPT_LSTM_MIN_CHUNK = 5
FTB_TUNE_LSTM_MIN_CHUNK = 1
FTB_TUNE_GRU_MIN_CHUNK = 2


class FtbError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f'ftb200 error {status}: {message}')
        self.status = status


class Tensor(C.Structure):
    _fields_ = [('name', C.c_char_p), ('data', C.c_void_p), ('dtype', C.c_int32), ('ndim', C.c_int32),
                ('shape', C.c_int64 * 4)]


class ConvDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('S', C.c_int32), ('Cin', C.c_int32), ('N', C.c_int32), ('ktaps', C.c_int32),
                ('pad_left', C.c_int32), ('lda', C.c_int32), ('ldo', C.c_int32), ('n_offset', C.c_int32),
                ('relu', C.c_int32), ('bias', C.c_void_p), ('scale', C.c_void_p), ('shift', C.c_void_p),
                ('residual_f32', C.c_void_p), ('residual_bf16', C.c_void_p), ('ldr', C.c_int32),
                ('out_scale', C.c_float), ('out_f32', C.c_void_p), ('out_bf16', C.c_void_p), ('out_t', C.c_void_p)]


class MelConfig(C.Structure):
    _fields_ = [('sample_rate', C.c_int32), ('n_fft', C.c_int32), ('hop_length', C.c_int32),
                ('win_length', C.c_int32), ('num_mels', C.c_int32), ('fmin', C.c_float), ('fmax', C.c_float)]


FT_INT_FIELDS = ['num_chars', 'embed_dims', 'series_embed_dims', 'durpred_conv_dims', 'durpred_rnn_dims',
                 'pitch_conv_dims', 'pitch_rnn_dims', 'energy_conv_dims', 'energy_rnn_dims', 'rnn_dims',
                 'prenet_dims', 'prenet_k', 'prenet_num_highways', 'postnet_dims', 'postnet_k',
                 'postnet_num_highways', 'n_mels']


class FtConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in FT_INT_FIELDS] + [('pitch_strength', C.c_float),
                                                          ('energy_strength', C.c_float), ('gemm_mode', C.c_int32)]


FP_INT_FIELDS = ['num_chars', 'n_mels', 'durpred_d_model', 'durpred_n_heads', 'durpred_layers', 'durpred_d_fft',
                 'pitch_d_model', 'pitch_n_heads', 'pitch_layers', 'pitch_d_fft', 'energy_d_model',
                 'energy_n_heads', 'energy_layers', 'energy_d_fft', 'd_model', 'conv1_kernel', 'conv2_kernel',
                 'prenet_layers', 'prenet_heads', 'prenet_fft', 'postnet_layers', 'postnet_heads', 'postnet_fft']


class FpConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in FP_INT_FIELDS] + [('pitch_strength', C.c_float),
                                                          ('energy_strength', C.c_float), ('gemm_mode', C.c_int32)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
# name -> (restype, argtypes): every symbol include/ftb200.h declares
SIGNATURES = {
    'ftb_last_error': (C.c_char_p, []),
    'ftb_abi_version': (_I, []),
    'ftb_struct_size': (_I, [_I]),
    'ftb_launch_count': (C.c_longlong, []),
    'ftb_profile_families': (_I, []),
    'ftb_profile_family_name': (C.c_char_p, [_I]),
    'ftb_profile_enable': (_I, [_I]),
    'ftb_profile_collect': (_I, [_P, _P, _P, _P]),
    'ftb_enable_peer_access': (_I, [_I, _I]),
    'ftb_ipc_alloc': (_I, [_L, _I, C.POINTER(_P), _P]),
    'ftb_ipc_open': (_I, [_P, _I, C.POINTER(_P)]),
    'ftb_ipc_release': (_I, [_P, _I]),
    'ftb_tune': (_I, [_I, _I]),
    'ftb_device_check': (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    'ftb_length_plan': (_I, [_P, _P, _P, _I, _I, _P]),
    'ftb_length_expand': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'ftb_length_index': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'ftb_duration_fallback': (_I, [_P, _L, _P, _P]),
    'ftb_zero_tail_rows': (_I, [_P, _I, _I, _L, _P, _P]),
    'ftb_duration_fallback_rows': (_I, [_P, _P, _I, _I, _P]),
    'ftb_conv_gemm_f32': (_I, [_P, _P, C.POINTER(ConvDesc), _P]),
    'ftb_conv_gemm_bf16': (_I, [_P, _P, C.POINTER(ConvDesc), _P]),
    'ftb_conv_bank_bf16': (_I, [_P, C.POINTER(_P), C.POINTER(ConvDesc), _I, _I, _P]),
    'ftb_tc_timeout_count': (_I, []),
    'ftb_pack_conv_weight': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'ftb_linear_pair': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    'ftb_rnn_bidir': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'ftb_rnn_bidir_rows': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'ftb_rnn_bidir_packed': (_I, [_P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'ftb_mel_create': (_I, [C.POINTER(MelConfig), _I, C.POINTER(_P)]),
    'ftb_mel_destroy': (None, [_P]),
    'ftb_mel_run': (_I, [_P, _P, _P, _P, _I, _L, _P, _I, _P]),
    'ftb_mel_filterbank': (_I, [_P, _P]),
    'ftb_mel_to_stft': (_I, [_P, _P, _I, _I, _I, _P, _P]),
    'ftb_griffinlim_workspace_bytes': (_L, [_I]),
    'ftb_griffinlim': (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _L, _P]),
    'ftb_trim_silence': (_I, [_P, _P, _I, _I, _F, _I, _I, _P, _P, _L, _P]),
    'ftb_ft_create': (_I, [C.POINTER(FtConfig), C.POINTER(Tensor), _I, _I, C.POINTER(_P)]),
    'ftb_ft_destroy': (None, [_P]),
    'ftb_ft_workspace_bytes': (_L, [_P, _I, _I, _I]),
    'ftb_ft_set_option': (_I, [_P, _I, _I]),
    'ftb_ft_predict': (_I, [_P, _P, _I, _I, _F, _P, _P, _P, _P, _L, _P]),
    'ftb_ft_synthesize': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _L, _P]),
    'ftb_ft_synthesize_packed': (_I, [_P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _P, _P, _P, _L, _P]),
    'ftb_ft_predict_ragged': (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _P, _P, _L, _P]),
    'ftb_ft_synthesize_ragged': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _L, _P]),
    'ftb_ft_series_predictor': (_I, [_P, _I, _P, _I, _I, _F, _P, _P, _L, _P]),
    'ftb_ft_cbhg': (_I, [_P, _I, _P, _I, _I, _P, _P, _L, _P]),
    'ftb_ft_last_launch_count': (_I, [_P]),
    'ftb_fp_create': (_I, [C.POINTER(FpConfig), C.POINTER(Tensor), _I, _I, C.POINTER(_P)]),
    'ftb_fp_destroy': (None, [_P]),
    'ftb_fp_workspace_bytes': (_L, [_P, _I, _I, _I]),
    'ftb_fp_predict': (_I, [_P, _P, _I, _I, _F, _P, _P, _P, _P, _L, _P]),
    'ftb_fp_synthesize': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _L, _P]),
    'ftb_fp_forward_eval': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _L, _P]),
    'ftb_fp_last_launch_count': (_I, [_P]),
    'ftb_attention_16': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
}
STRUCT_IDS = {0: Tensor, 1: ConvDesc, 2: MelConfig, 3: FtConfig, 4: FpConfig}


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load (building first if needed) csrc/libftb200.so and set the prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not _LIB_PATH.exists():
        from .build import build  # raises if nvcc is unavailable: there is no fallback
        build()
    handle = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if handle.ftb_abi_version() != 1:
        raise RuntimeError('libftb200.so ABI version mismatch; rebuild with forwardtacotron_b200/build.py')
    for sid, cls in STRUCT_IDS.items():
        if handle.ftb_struct_size(sid) != C.sizeof(cls):
            raise RuntimeError(f'ctypes layout of {cls.__name__} does not match include/ftb200.h')
    _LIB = handle
    return handle


def check(status: int) -> None:
    if status != 0:
        raise FtbError(status, lib().ftb_last_error().decode('utf-8', 'replace'))


def current_stream(device) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def tensor_table(named: Iterable[Tuple[str, 'object']]) -> Tuple[C.Array, List[object]]:
    """state_dict items -> (ftb_tensor array, keep-alive list).  Tensors must live on the GPU."""
    import torch
    dt = {torch.float32: FTB_F32, torch.int64: FTB_I64, torch.bfloat16: FTB_BF16, torch.int32: FTB_I32}
    items = [(k, v) for k, v in named]
    arr = (Tensor * len(items))()
    keep: List[object] = []
    for i, (k, v) in enumerate(items):
        if v.dtype not in dt:
            raise TypeError(f'state_dict entry {k} has unsupported dtype {v.dtype}')
        if v.dim() > 4:
            raise ValueError(f'state_dict entry {k} has more than 4 dims')
        v = v.contiguous()
        name = k.encode()
        keep += [v, name]
        arr[i].name = name
        arr[i].data = v.data_ptr()
        arr[i].dtype = dt[v.dtype]
        arr[i].ndim = v.dim()
        for j, s in enumerate(v.shape):
            arr[i].shape[j] = s
    return arr, keep
