"""Spectrogram representations around the STFT pair, one fused pass each.

``split(X, kind)`` / ``join(a, b, kind)`` with ``kind`` in

* ``'real_imag'``        -- ``STFT.forward(return_type='real_imag')`` / ``input_type`` (stft.py:91-110)
* ``'mag_phase'``        -- ``spec.abs(), spec.angle()`` / ``torch.polar`` (stft.py:95-99,106-108)
* ``'log1p_mag_phase'``  -- MetricGAN-OKD's wrappers (metricganokd.py:185-195):
  ``log1p(|X| + eps), angle X`` and ``expm1(mag) * exp(1j * phase)``

Each is ONE elementwise kernel over the tensors' memory (``brv_spec_split`` / ``brv_spec_join``;
the reference runs 2-4 eager passes) and differentiable (``brv_spec_*_grad``).  Outputs keep the
strides of the input, like the elementwise torch ops they replace: the planes of a frame-major
spectrogram come back as ``(..., F, T)`` views with strides ``(1, F)``.
"""
import torch

from .. import _lib

_MODES = {'real_imag': 0, 'mag_phase': 1, 'log1p_mag_phase': 2}


def _mode(kind):
    try:
        return _MODES[kind]
    except KeyError:
        raise ValueError(f'kind must be one of {sorted(_MODES)}, got {kind}') from None


def _is_dense(t):
    """True when the elements of `t` fill one gap-free block of memory (in any dim order)."""
    expected = 1
    for stride, size in sorted((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1):
        if stride != expected:
            return False
        expected *= size
    return True


def _dense(t):
    """`t` itself when its memory is dense (any dim order), else a contiguous copy."""
    return t if _is_dense(t) else t.contiguous()


def _like(t, ref, dtype):
    """`t` as `dtype` with exactly the strides of `ref` (copy only if they differ)."""
    if t.dtype == dtype and t.stride() == ref.stride():
        return t
    out = torch.empty_strided(ref.shape, ref.stride(), dtype=dtype, device=ref.device)
    out.copy_(t)
    return out


def _split_raw(X, mode, eps):
    a = torch.empty_strided(X.shape, X.stride(), dtype=torch.float32, device=X.device)
    b = torch.empty_strided(X.shape, X.stride(), dtype=torch.float32, device=X.device)
    with _lib.on_device(X.device):
        _lib.check(_lib.lib().brv_spec_split(_lib.ptr(X), X.numel(), mode, float(eps), _lib.ptr(a),
                                             _lib.ptr(b), _lib.stream_ptr(X.device)))
    return a, b


def _join_raw(a, b, mode):
    X = torch.empty_strided(a.shape, a.stride(), dtype=torch.complex64, device=a.device)
    with _lib.on_device(a.device):
        _lib.check(_lib.lib().brv_spec_join(_lib.ptr(a), _lib.ptr(b), a.numel(), mode, _lib.ptr(X),
                                            _lib.stream_ptr(a.device)))
    return X


class _SplitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, mode, eps):
        ctx.mode, ctx.eps = mode, eps
        ctx.save_for_backward(X)
        return _split_raw(X, mode, eps)

    @staticmethod
    def backward(ctx, ga, gb):
        X, = ctx.saved_tensors
        ga = None if ga is None else _like(ga, X, torch.float32)
        gb = None if gb is None else _like(gb, X, torch.float32)
        gX = torch.empty_strided(X.shape, X.stride(), dtype=torch.complex64, device=X.device)
        with _lib.on_device(X.device):
            _lib.check(_lib.lib().brv_spec_split_grad(
                _lib.ptr(ga), _lib.ptr(gb), _lib.ptr(X), X.numel(), ctx.mode, float(ctx.eps),
                _lib.ptr(gX), _lib.stream_ptr(X.device)))
        return gX, None, None


class _JoinFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, mode):
        ctx.mode = mode
        ctx.save_for_backward(a, b)
        return _join_raw(a, b, mode)

    @staticmethod
    def backward(ctx, gX):
        a, b = ctx.saved_tensors
        gX = _like(gX.resolve_conj().resolve_neg(), a, torch.complex64)
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        with _lib.on_device(a.device):
            _lib.check(_lib.lib().brv_spec_join_grad(
                _lib.ptr(gX), _lib.ptr(a), _lib.ptr(b), a.numel(), ctx.mode, _lib.ptr(ga),
                _lib.ptr(gb), _lib.stream_ptr(a.device)))
        return ga, gb, None


def split(X, kind='mag_phase', eps=0.0):
    """complex64 CUDA spectrogram -> two float32 tensors of the same shape and strides."""
    mode = _mode(kind)
    _lib.require_cuda(X, 'spectrogram')
    if X.dtype != torch.complex64:
        raise RuntimeError(f'spectrogram must be complex64, got {X.dtype}')
    X = _dense(X.resolve_conj().resolve_neg())
    if torch.is_grad_enabled() and X.requires_grad:
        return _SplitFunction.apply(X, mode, eps)
    return _split_raw(X, mode, eps)


def join(a, b, kind='mag_phase'):
    """Two real CUDA tensors of one shape -> complex64 spectrogram (strides of `a`)."""
    mode = _mode(kind)
    _lib.require_cuda(a, 'spectrogram plane')
    _lib.require_cuda(b, 'spectrogram plane')
    if a.shape != b.shape:
        raise RuntimeError(f'planes must have the same shape, got {tuple(a.shape)} and {tuple(b.shape)}')
    a = _dense(a.float())
    b = _like(b, a, torch.float32)
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return _JoinFunction.apply(a, b, mode)
    return _join_raw(a, b, mode)
