"""ctypes binding of the C-ABI library (include/brever_b200.h).

The shared object is built in-tree by ``__graft_entry__.build()`` (plain
``nvcc -shared``; no torch extension machinery) and loaded here.  There is no
CPU or pure-PyTorch fallback anywhere in this package: if the library is
missing, or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbrever_b200.so')

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOLA, ERR_ALLOC = 0, -1, -2, -3, -4, -5

_c = ctypes
_i64, _int, _f32, _f64, _ptr, _sz = (_c.c_int64, _c.c_int, _c.c_float,
                                      _c.c_double, _c.c_void_p, _c.c_size_t)

# name -> (restype, argtypes); mirrors include/brever_b200.h one to one
PROTOTYPES = {
    'brv_abi_version': (_int, []),
    'brv_status_string': (_c.c_char_p, [_int]),
    'brv_last_error': (_c.c_char_p, []),
    'brv_launch_count': (_c.c_uint64, []),
    'brv_set_force_generic': (_int, [_int]),
    'brv_set_tc_variant': (_int, [_int]),
    'brv_device_query': (_int, [_ptr, _ptr, _ptr]),
    'brv_stft_plan_create': (_int, [_ptr, _int, _int, _int, _ptr, _int, _int,
                                    _f64, _f64]),
    'brv_stft_plan_destroy': (_int, [_ptr]),
    'brv_stft_geometry': (_int, [_ptr, _i64, _ptr, _ptr, _ptr]),
    'brv_istft_geometry': (_int, [_ptr, _i64, _ptr]),
    'brv_stft_forward': (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr]),
    'brv_stft_forward_grad': (_int, [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64,
                                     _ptr, _ptr, _sz, _ptr]),
    'brv_istft_forward': (_int, [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64,
                                 _ptr, _ptr, _sz, _ptr]),
    'brv_istft_forward_grad': (_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _sz,
                                      _ptr]),
    'brv_stft_forward_f64': (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr]),
    'brv_istft_forward_f64': (_int, [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr]),
    'brv_stft_forward_grad_f64': (_int, [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr]),
    'brv_istft_forward_grad_f64': (_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr]),
    'brv_stft_workspace_bytes': (_sz, [_ptr, _i64, _i64]),
    'brv_stft_workspace_bytes_op': (_sz, [_ptr, _i64, _i64, _int]),
    'brv_convstft_geometry': (_int, [_ptr, _i64, _ptr]),
    'brv_convstft_forward': (_int, [_ptr, _ptr, _i64, _i64, _i64, _int, _ptr, _ptr]),
    'brv_convstft_backward': (_int, [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _int,
                                     _ptr, _ptr]),
    'brv_mel_apply': (_int, [_ptr, _i64, _i64, _i64, _i64, _int, _i64, _ptr,
                             _ptr, _ptr, _int, _ptr, _ptr]),
    'brv_fbe_features': (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _int, _int,
                                _i64, _ptr, _ptr, _ptr, _int, _int, _int, _int,
                                _f32, _int, _int, _ptr, _ptr, _ptr, _ptr]),
    'brv_mel_features': (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _int, _int,
                                _i64, _int, _ptr, _ptr, _ptr, _int, _int, _int,
                                _int, _f32, _ptr, _int, _ptr, _ptr]),
    'brv_ic_coherence': (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _int, _int,
                                _i64, _f32, _f32, _ptr, _ptr]),
    'brv_stack_normalize': (_int, [_ptr, _i64, _int, _i64, _int, _int, _ptr,
                                   _ptr, _ptr, _ptr]),
    'brv_cumulative_normalize': (_int, [_ptr, _i64, _i64, _f32, _ptr, _ptr]),
    'brv_snr_forward': (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64,
                               _i64, _i64, _int, _f32, _f32, _ptr, _ptr, _ptr,
                               _sz, _ptr]),
    'brv_snr_workspace_bytes': (_sz, [_i64, _i64]),
    'brv_masked_affine': (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64,
                                 _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr,
                                 _ptr, _ptr]),
    'brv_apply_mask': (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr]),
    'brv_criterion_backward': (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _f32, _ptr, _int,
                                      _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _ptr,
                                      _ptr]),
    'brv_l1_workspace_bytes': (_sz, [_i64, _i64]),
    'brv_l1_forward': (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64,
                              _i64, _i64, _ptr, _ptr, _sz, _ptr]),
    'brv_l1_backward': (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64,
                               _i64, _i64, _i64, _ptr, _ptr]),
    'brv_mag_l1_forward': (_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _sz, _ptr]),
    'brv_mag_l1_backward': (_int, [_ptr, _ptr, _ptr, _i64, _i64, _ptr, _ptr]),
    'brv_stft_plan_set_framing': (_int, [_ptr, _int, _int]),
    'brv_reflect_pad': (_int, [_ptr, _i64, _i64, _i64, _i64, _int, _int, _ptr, _ptr]),
    'brv_reflect_pad_grad': (_int, [_ptr, _i64, _i64, _i64, _int, _int, _ptr, _ptr]),
    'brv_mrstft_forward': (_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr]),
    'brv_mrstft_backward': (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _ptr, _ptr]),
    'brv_channel_mean_mask': (_int, [_ptr, _i64, _i64, _i64, _i64, _ptr, _i64, _i64, _i64, _i64, _int, _int,
                                     _i64, _ptr, _ptr]),
    'brv_accumulate_mean': (_int, [_ptr, _i64, _ptr, _ptr]),
    'brv_spec_split': (_int, [_ptr, _i64, _int, _f32, _ptr, _ptr, _ptr]),
    'brv_spec_join': (_int, [_ptr, _ptr, _i64, _int, _ptr, _ptr]),
    'brv_spec_split_grad': (_int, [_ptr, _ptr, _ptr, _i64, _int, _f32, _ptr, _ptr]),
    'brv_spec_join_grad': (_int, [_ptr, _ptr, _ptr, _i64, _int, _ptr, _ptr, _ptr]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raise if it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with '
                '`python -c "import __graft_entry__ as g; g.build()"`. '
                'brever_b200 has no CPU or PyTorch fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status):
    """Map a brv_status to the exception the reference would raise."""
    if status == OK:
        return
    msg = lib().brv_last_error().decode() or \
        lib().brv_status_string(status).decode()
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)  # CUDA errors, NOLA (torch.istft raises RuntimeError)


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{what} must be a torch.Tensor, got {type(t)}')
    if not t.is_cuda:
        raise RuntimeError(
            f'{what} is on {t.device}: brever_b200 runs on CUDA tensors only '
            '(there is deliberately no CPU fallback; move the tensor to the '
            'GPU or keep using brever.modules on CPU)')


class _NullContext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullContext()


def on_device(device):
    """Context that makes `device` current for a launch; free when it already is."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
