"""Drop-in ``STFT`` and ``MelFilterbank`` backed by the sm_100a kernels.

Same constructor keywords, attributes and call signatures as
``brever/modules/stft.py:12-198``; the arithmetic runs in
``libbrever_b200.so`` (include/brever_b200.h).  The objects stay plain Python
(not ``nn.Module``), hold no ``state_dict`` entries and are picklable, like the
reference's; device-side constants are created lazily per CUDA device.

Deliberate differences:
  * CUDA tensors only — a CPU tensor raises instead of silently running torch;
  * ``backward`` never modifies its input (the reference divides a complex
    input by ``scale_factor`` in place, stft.py:114);
  * ``pad_mode`` ``'constant'`` and ``'reflect'``; ``center=False`` for the forward transform
    and its gradient only (the reference's own docstring says to always use ``center=True``).
"""
import ctypes
import functools
import math

import numpy as np
import scipy.signal
import torch

from .. import _lib
from . import specfmt


def _fft_freqs(fs, n_fft):
    # brever/utils.py:40-66, onesided
    freqs = np.arange(n_fft) * fs / n_fft
    return freqs[freqs <= fs / 2]


class _StftFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, stft):
        ctx.stft = stft
        ctx.samples = x2d.shape[-1]
        return stft._forward_raw(x2d)

    @staticmethod
    def backward(ctx, grad):
        return ctx.stft._forward_grad_raw(grad, ctx.samples), None


class _ReflectPadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, padded, left, right_reflect):
        ctx.geom = (x2d.shape[-1], padded, left, right_reflect)
        return _reflect_pad_raw(x2d, padded, left, right_reflect)

    @staticmethod
    def backward(ctx, grad):
        samples, padded, left, right_reflect = ctx.geom
        grad = grad.float().contiguous()
        gx = torch.empty((grad.shape[0], samples), dtype=torch.float32, device=grad.device)
        with _lib.on_device(grad.device):
            _lib.check(_lib.lib().brv_reflect_pad_grad(
                _lib.ptr(grad), grad.shape[0], samples, padded, left, int(right_reflect),
                _lib.ptr(gx), _lib.stream_ptr(grad.device)))
        return gx, None, None, None


def _reflect_pad_raw(x2d, padded, left, right_reflect):
    n_sig, samples = x2d.shape
    if x2d.stride(-1) != 1:
        x2d = x2d.contiguous()
    if left >= padded > 0 or (right_reflect and padded - samples >= samples > 0):
        raise RuntimeError('Padding size should be less than the corresponding input dimension')
    out = torch.empty((n_sig, padded + 2 * left), dtype=torch.float32, device=x2d.device)
    with _lib.on_device(x2d.device):
        _lib.check(_lib.lib().brv_reflect_pad(
            _lib.ptr(x2d), n_sig, samples, x2d.stride(0) if n_sig > 1 else samples, padded, left,
            int(right_reflect), _lib.ptr(out), _lib.stream_ptr(x2d.device)))
    return out


def _reflect_pad(x2d, padded, left, right_reflect):
    if torch.is_grad_enabled() and x2d.requires_grad:
        return _ReflectPadFunction.apply(x2d, padded, left, right_reflect)
    return _reflect_pad_raw(x2d, padded, left, right_reflect)


class _IstftFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec3d, stft):
        ctx.stft = stft
        ctx.frames = spec3d.shape[-1]
        ctx.in_dtype = spec3d.dtype
        return stft._inverse_raw(spec3d)

    @staticmethod
    def backward(ctx, grad):
        g = ctx.stft._inverse_grad_raw(grad, ctx.frames)
        return g.to(ctx.in_dtype), None


class _StftFunction64(torch.autograd.Function):
    """float64 tensors: the double-precision kernels (brv_stft_f64.cu)."""
    @staticmethod
    def forward(ctx, x2d, stft):
        ctx.stft = stft
        ctx.samples = x2d.shape[-1]
        return stft._forward_raw64(x2d)

    @staticmethod
    def backward(ctx, grad):
        return ctx.stft._forward_grad_raw64(grad, ctx.samples), None


class _IstftFunction64(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec3d, stft):
        ctx.stft = stft
        ctx.frames = spec3d.shape[-1]
        return stft._inverse_raw64(spec3d)

    @staticmethod
    def backward(ctx, grad):
        return ctx.stft._inverse_grad_raw64(grad, ctx.frames), None


class STFT:
    """Short-time Fourier transform with the reference's conventions.

    Right-pads to whole frames, centre-pads by ``n_fft // 2``, periodic window
    from scipy, energy normalisation by ``sqrt(sum w^2)``, optional magnitude
    compression ``|X|^c e^{j angle X}`` and scale, arbitrary leading dims
    (brever/modules/stft.py:32-149).
    """

    def __init__(self, frame_length=512, hop_length=256, window='hann',
                 center=True, pad_mode='constant', normalized=True,
                 onesided=True, compression_factor=1, scale_factor=1,
                 n_fft=None):
        if pad_mode not in ('constant', 'reflect'):
            raise NotImplementedError(
                "brever_b200.STFT supports pad_mode='constant' and 'reflect' "
                f'(got {pad_mode!r}; every reference model uses constant)')
        # raw torch.stft framing (no STFT.pad to whole frames): MANNER's loss (manner.py)
        self._raw_framing = False
        self.frame_length = frame_length
        self.hop_length = hop_length
        self.center = center
        self.pad_mode = pad_mode
        self.normalized = normalized
        self.onesided = onesided
        self.compression_factor = compression_factor
        self.scale_factor = scale_factor
        self.n_fft = frame_length if n_fft is None else n_fft

        # same window resolution order as stft.py:46-54
        if window is None:
            window = 'boxcar'
        if isinstance(window, str):
            window = functools.partial(scipy.signal.get_window, window)
        if callable(window):
            window = window(frame_length)
        if isinstance(window, np.ndarray):
            window = torch.from_numpy(window)
        self.window = window
        self._plans = {}

    # -- pickling: plans are device handles, rebuild lazily ------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_plans'] = {}
        return state

    def __del__(self):
        try:
            for plan in self._plans.values():
                _lib.lib().brv_stft_plan_destroy(plan)
        except Exception:
            pass

    def _plan(self, device, inverse=False):
        """Device plan; the forward plan carries the framing of this object (centre / reflect
        padding / raw framing), the inverse plan is always the centred one torch.istft needs."""
        index = device.index if device.index is not None else torch.cuda.current_device()
        # reflect padding is materialised by brv_reflect_pad: the kernels then see plain frames
        center = bool(self.center) and self.pad_mode == 'constant'
        pad_frames = self.pad_mode == 'constant' and not self._raw_framing
        default = center and pad_frames
        key = index if (inverse or default) else (index, 'framed')
        plan = self._plans.get(key)
        if plan is None:
            win = self.window.detach().to('cpu', torch.float64).contiguous()
            if win.numel() != self.frame_length:
                raise ValueError('window must have frame_length samples, got '
                                 f'{win.numel()}')
            handle = ctypes.c_void_p()
            with torch.cuda.device(key):
                _lib.check(_lib.lib().brv_stft_plan_create(
                    ctypes.byref(handle), int(self.frame_length),
                    int(self.hop_length), int(self.n_fft),
                    ctypes.c_void_p(win.data_ptr()), int(bool(self.normalized)),
                    int(bool(self.onesided)), float(self.compression_factor),
                    float(self.scale_factor)))
                if key != index:
                    _lib.check(_lib.lib().brv_stft_plan_set_framing(handle, int(center), int(pad_frames)))
            plan = self._plans[key] = handle
        return plan

    # -- integer frame arithmetic (stft.py:140-149) ---------------------------
    def frame_count(self, samples):
        # frames WITHOUT the n_fft//2 centre padding torch.stft adds
        return math.ceil(max(samples - self.frame_length, 0)
                         / self.hop_length) + 1

    def pad(self, x):
        samples = x.shape[-1]
        padding = (self.frame_count(samples) - 1) * self.hop_length \
            + self.frame_length - samples
        return torch.nn.functional.pad(x, (0, padding), mode=self.pad_mode)

    def _right_pad(self, samples):
        if self._raw_framing:
            return 0
        return (self.frame_count(samples) - 1) * self.hop_length + self.frame_length - samples

    def n_frames(self, samples):
        """Frames the forward transform returns for `samples` input samples."""
        padded = samples + self._right_pad(samples) + (2 * (self.n_fft // 2) if self.center else 0)
        return 1 + (padded - self.n_fft) // self.hop_length

    @property
    def n_bins(self):
        return self.n_fft // 2 + 1 if self.onesided else self.n_fft

    # -- raw kernels on flattened tensors -------------------------------------
    def _forward_raw(self, x2d):
        """(n_sig, S) float32 CUDA -> (n_sig, F, T) complex64 view of (n_sig, T, F)."""
        n_sig, samples = x2d.shape
        if x2d.stride(-1) != 1:
            x2d = x2d.contiguous()
        frames = self.n_frames(samples) if self.pad_mode == 'constant' else \
            1 + (samples - self.n_fft) // self.hop_length      # reflect: x2d is already padded
        out = torch.empty((n_sig, frames, self.n_bins), dtype=torch.complex64,
                          device=x2d.device)
        if n_sig:
            with _lib.on_device(x2d.device):
                _lib.check(_lib.lib().brv_stft_forward(
                    self._plan(x2d.device), _lib.ptr(x2d), n_sig, samples,
                    x2d.stride(0) if n_sig > 1 else samples, _lib.ptr(out),
                    _lib.stream_ptr(x2d.device)))
        return out.transpose(1, 2)

    def _forward_grad_raw(self, grad, samples):
        n_sig, bins, frames = grad.shape
        grad = grad.to(torch.complex64).resolve_conj().resolve_neg()   # raw pointers below: no lazy bits
        gx = torch.empty((n_sig, samples), dtype=torch.float32, device=grad.device)
        if n_sig:
            lib = _lib.lib()
            with _lib.on_device(grad.device):
                plan = self._plan(grad.device)
                nbytes = lib.brv_stft_workspace_bytes_op(plan, n_sig, frames, 1)
                ws = _lib.workspace(nbytes, grad.device)
                _lib.check(lib.brv_stft_forward_grad(
                    plan, _lib.ptr(grad), grad.stride(0), grad.stride(1),
                    grad.stride(2), n_sig, samples, _lib.ptr(gx), _lib.ptr(ws),
                    nbytes, _lib.stream_ptr(grad.device)))
        return gx

    def _inverse_raw(self, spec3d):
        """(n_sig, F, T) complex (any strides) -> (n_sig, hop*(T-1)) float32."""
        n_sig, bins, frames = spec3d.shape
        if bins != self.n_bins:
            raise RuntimeError(f'expected {self.n_bins} frequency bins, got {bins}')
        spec3d = spec3d.to(torch.complex64).resolve_conj().resolve_neg()   # raw pointers below
        out_len = self.hop_length * (frames - 1) + self.n_fft - 2 * (self.n_fft // 2)
        y = torch.empty((n_sig, out_len), dtype=torch.float32, device=spec3d.device)
        lib = _lib.lib()
        if not self.center:
            raise NotImplementedError('STFT.backward is implemented for center=True only')
        with _lib.on_device(spec3d.device):
            plan = self._plan(spec3d.device, inverse=True)
            nbytes = lib.brv_stft_workspace_bytes_op(plan, n_sig, frames, 0)
            ws = _lib.workspace(nbytes, spec3d.device)
            _lib.check(lib.brv_istft_forward(
                plan, _lib.ptr(spec3d), spec3d.stride(0), spec3d.stride(1),
                spec3d.stride(2), n_sig, frames, _lib.ptr(y), _lib.ptr(ws),
                nbytes, _lib.stream_ptr(spec3d.device)))
        return y

    def _inverse_grad_raw(self, grad, frames):
        n_sig = grad.shape[0]
        grad = grad.to(torch.float32).contiguous()
        gX = torch.empty((n_sig, frames, self.n_bins), dtype=torch.complex64,
                         device=grad.device)
        if n_sig:
            lib = _lib.lib()
            with _lib.on_device(grad.device):
                plan = self._plan(grad.device, inverse=True)
                nbytes = lib.brv_stft_workspace_bytes_op(plan, n_sig, frames, 2)
                ws = _lib.workspace(nbytes, grad.device)
                _lib.check(lib.brv_istft_forward_grad(
                    plan, _lib.ptr(grad), n_sig, frames, _lib.ptr(gX),
                    _lib.ptr(ws), nbytes, _lib.stream_ptr(grad.device)))
        return gX.transpose(1, 2)

    # -- float64 tensors: double-precision kernels (torch.stft / istft compute in the input's
    #    precision; the tensor-core path is fp32-grade) -----------------------------------------
    def _forward_raw64(self, x2d):
        n_sig, samples = x2d.shape
        if x2d.stride(-1) != 1:
            x2d = x2d.contiguous()
        frames = self.n_frames(samples) if self.pad_mode == 'constant' else \
            1 + (samples - self.n_fft) // self.hop_length      # reflect: x2d is already padded
        out = torch.empty((n_sig, frames, self.n_bins), dtype=torch.complex128, device=x2d.device)
        if n_sig:
            with _lib.on_device(x2d.device):
                _lib.check(_lib.lib().brv_stft_forward_f64(
                    self._plan(x2d.device), _lib.ptr(x2d), n_sig, samples,
                    x2d.stride(0) if n_sig > 1 else samples, _lib.ptr(out), _lib.stream_ptr(x2d.device)))
        return out.transpose(1, 2)

    def _forward_grad_raw64(self, grad, samples):
        n_sig, bins, frames = grad.shape
        grad = grad.to(torch.complex128).resolve_conj().resolve_neg()
        gx = torch.empty((n_sig, samples), dtype=torch.float64, device=grad.device)
        if n_sig:
            with _lib.on_device(grad.device):
                _lib.check(_lib.lib().brv_stft_forward_grad_f64(
                    self._plan(grad.device), _lib.ptr(grad), grad.stride(0), grad.stride(1), grad.stride(2),
                    n_sig, samples, _lib.ptr(gx), _lib.stream_ptr(grad.device)))
        return gx

    def _inverse_raw64(self, spec3d):
        n_sig, bins, frames = spec3d.shape
        if bins != self.n_bins:
            raise RuntimeError(f'expected {self.n_bins} frequency bins, got {bins}')
        if not self.center:
            raise NotImplementedError('STFT.backward is implemented for center=True only')
        spec3d = spec3d.resolve_conj().resolve_neg()
        out_len = self.hop_length * (frames - 1) + self.n_fft - 2 * (self.n_fft // 2)
        y = torch.empty((n_sig, out_len), dtype=torch.float64, device=spec3d.device)
        with _lib.on_device(spec3d.device):
            _lib.check(_lib.lib().brv_istft_forward_f64(
                self._plan(spec3d.device, inverse=True), _lib.ptr(spec3d), spec3d.stride(0), spec3d.stride(1),
                spec3d.stride(2), n_sig, frames, _lib.ptr(y), _lib.stream_ptr(spec3d.device)))
        return y

    def _inverse_grad_raw64(self, grad, frames):
        n_sig = grad.shape[0]
        grad = grad.to(torch.float64).contiguous()
        gX = torch.empty((n_sig, frames, self.n_bins), dtype=torch.complex128, device=grad.device)
        if n_sig:
            with _lib.on_device(grad.device):
                _lib.check(_lib.lib().brv_istft_forward_grad_f64(
                    self._plan(grad.device, inverse=True), _lib.ptr(grad), n_sig, frames, _lib.ptr(gX),
                    _lib.stream_ptr(grad.device)))
        return gX.transpose(1, 2)

    # -- public API (stft.py:56-138) -------------------------------------------
    def __call__(self, x, return_type='complex'):
        return self.forward(x, return_type=return_type)

    def forward(self, x, return_type='complex'):
        if return_type not in ('complex', 'real_imag', 'mag_phase'):
            raise ValueError('return_type must be complex, real_imag or '
                             f'mag_phase, got {return_type}')
        _lib.require_cuda(x, 'STFT input')
        if x.is_complex():
            raise RuntimeError('STFT input must be real')
        in_dtype = x.dtype
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        if in_dtype == torch.float64:
            # double-precision kernels; reflect padding by torch (a fidelity path, not a fast one)
            if self.pad_mode == 'reflect':
                pad = torch.nn.functional.pad
                right = 0 if self._raw_framing else self._right_pad(x2d.shape[-1])
                if right:
                    x2d = pad(x2d.unsqueeze(1), (0, right), mode='reflect').squeeze(1)
                if self.center:
                    x2d = pad(x2d.unsqueeze(1), (self.n_fft // 2, self.n_fft // 2), mode='reflect').squeeze(1)
            if torch.is_grad_enabled() and x2d.requires_grad:
                spec = _StftFunction64.apply(x2d, self)
            else:
                spec = self._forward_raw64(x2d)
            spec = spec.view(*lead, *spec.shape[-2:])
            if return_type == 'complex':
                return spec
            return (spec.real, spec.imag) if return_type == 'real_imag' else (spec.abs(), spec.angle())
        if x2d.dtype != torch.float32:
            x2d = x2d.float()  # fp16/bf16 (AMP) and fp64 compute in fp32
        if self.pad_mode == 'reflect':
            # STFT.pad mirrors the tail (F.pad(mode=pad_mode), stft.py:140-144), torch.stft the
            # centre padding: one gather kernel materialises both
            x2d = _reflect_pad(x2d, x2d.shape[-1] + self._right_pad(x2d.shape[-1]),
                               self.n_fft // 2 if self.center else 0, not self._raw_framing)
        if torch.is_grad_enabled() and x2d.requires_grad:
            spec = _StftFunction.apply(x2d, self)
        else:
            spec = self._forward_raw(x2d)
        spec = spec.view(*lead, *spec.shape[-2:])
        if in_dtype == torch.float64:
            spec = spec.to(torch.complex128)
        if return_type == 'complex':
            return spec
        if spec.dtype != torch.complex64:          # float64 callers: torch ops on the cast result
            return (spec.real, spec.imag) if return_type == 'real_imag' else (spec.abs(), spec.angle())
        # one fused pass (brv_spec_split) instead of the reference's eager .abs() / .angle()
        return specfmt.split(spec, return_type)

    def backward(self, x, input_type='complex'):
        if input_type in ('real_imag', 'mag_phase'):
            a, b = x
            _lib.require_cuda(a, 'STFT.backward input')
            if a.dtype == torch.float64:
                x = torch.complex(a, b) if input_type == 'real_imag' else torch.polar(a, b)
            else:
                x = specfmt.join(a, b, input_type)     # one fused pass (brv_spec_join)
        elif input_type != 'complex':
            raise ValueError('input_type must be complex, real_imag or '
                             f'mag_phase, got {input_type}')
        _lib.require_cuda(x, 'STFT.backward input')
        if not x.is_complex():
            raise RuntimeError('STFT.backward input must be complex')
        out_dtype = {torch.complex32: torch.float16,
                     torch.complex128: torch.float64}.get(x.dtype, torch.float32)
        lead = x.shape[:-2]
        spec3d = x.reshape(-1, *x.shape[-2:])
        if x.dtype == torch.complex128:
            if torch.is_grad_enabled() and spec3d.requires_grad:
                y = _IstftFunction64.apply(spec3d, self)
            else:
                y = self._inverse_raw64(spec3d)
            return y.view(*lead, -1)
        if torch.is_grad_enabled() and spec3d.requires_grad:
            y = _IstftFunction.apply(spec3d, self)
        else:
            y = self._inverse_raw(spec3d)
        y = y.view(*lead, -1)
        return y if out_dtype == torch.float32 else y.to(out_dtype)


class ConvSTFT:
    """Drop-in ``ConvSTFT`` (brever/modules/stft.py:201-319): the STFT written as a strided
    convolution with windowed DFT rows, and its transposed convolution as the synthesis.

    Same constructor keywords, attributes (``window`` is the square root of the scipy window,
    ``filters`` the ``(2F, 1, L)`` float32 kernel bank, ``_normalization_factor``), padding
    (``pad``, ``frame_count``) and return / input types as the reference.  The arithmetic is
    the folded tensor-core pair of ``brv_convstft_forward`` / ``brv_convstft_backward``:
    frames start ``L - H`` samples before ``t * H``, the DC row carries ``1 / sqrt(2)``, the
    synthesis is the exact adjoint (no envelope division) trimmed by ``L - H`` per side.
    ``frame_length`` in {128, 256, 384, 512} with ``hop_length`` = L/4, L/2 or L runs on the folded
    tensor-core kernels, every other size on the direct-sum kernels of ``brv_stft_f64.cu`` (correct,
    not fast); differentiable in both directions through the adjoint pair
    (``compression_factor == 1``).
    """

    def __init__(self, frame_length=512, hop_length=256, window='hann',
                 compression_factor=1, scale_factor=1, normalized=True):
        self.frame_length = frame_length
        self.hop_length = hop_length
        self.compression_factor = compression_factor
        self.scale_factor = scale_factor
        self.normalized = normalized
        if isinstance(window, str):
            window = scipy.signal.get_window(window, frame_length) ** 0.5   # stft.py:213-214
        if isinstance(window, np.ndarray):
            window = torch.from_numpy(window)
        self.window = window
        self._normalization_factor = 0.5 * frame_length / hop_length ** 0.5   # stft.py:232
        # the kernels apply the normalisation themselves: the core plan is un-normalised
        self._core = STFT(frame_length=frame_length, hop_length=hop_length, window=window,
                          normalized=False, compression_factor=compression_factor,
                          scale_factor=scale_factor)

    @property
    def filters(self):
        """The reference's convolution kernels (stft.py:218-238), for inspection."""
        L = self.frame_length
        filters = torch.fft.fft(torch.eye(L))[:L // 2 + 1]
        filters[0, :] /= 2 ** 0.5
        if self.normalized:
            filters /= self._normalization_factor
        filters *= self.window
        return torch.cat([filters.real, filters.imag]).unsqueeze(1).float()

    def __call__(self, x, return_type='complex'):
        return self.forward(x, return_type=return_type)

    def frame_count(self, samples):
        return math.ceil(max(samples - self.frame_length, 0) / self.hop_length) + 1

    def pad(self, x):
        samples = x.shape[-1]
        padding = (self.frame_count(samples) - 1) * self.hop_length + self.frame_length - samples
        x = torch.nn.functional.pad(x, (0, padding))
        padding = self.frame_length - self.hop_length
        return torch.nn.functional.pad(x, (padding, padding))

    def n_frames(self, samples):
        padded = (self.frame_count(samples) - 1) * self.hop_length + self.frame_length \
            + 2 * (self.frame_length - self.hop_length)
        return (padded - self.frame_length) // self.hop_length + 1

    def forward(self, x, return_type='complex'):
        if return_type not in ('complex', 'real_imag', 'mag_phase'):
            raise ValueError('return_type must be complex, real_imag or '
                             f'mag_phase, got {return_type}')
        _lib.require_cuda(x, 'ConvSTFT input')
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        if x2d.dtype != torch.float32:
            x2d = x2d.float()
        if x2d.stride(-1) != 1:
            x2d = x2d.contiguous()
        bins = self.frame_length // 2 + 1
        if torch.is_grad_enabled() and x2d.requires_grad:
            spec = _ConvForwardFunction.apply(x2d, self)
        else:
            spec = self._conv_forward_raw(x2d)
        spec = spec.view(*lead, bins, spec.shape[-1])
        if return_type == 'complex':
            return spec
        return specfmt.split(spec, return_type)

    def _conv_forward_raw(self, x2d):
        n_sig, samples = x2d.shape
        frames, bins = self.n_frames(samples), self.frame_length // 2 + 1
        out = torch.empty((n_sig, frames, bins), dtype=torch.complex64, device=x2d.device)
        if n_sig:
            with _lib.on_device(x2d.device):
                _lib.check(_lib.lib().brv_convstft_forward(
                    self._core._plan(x2d.device), _lib.ptr(x2d), n_sig, samples,
                    x2d.stride(0) if n_sig > 1 else samples, int(bool(self.normalized)),
                    _lib.ptr(out), _lib.stream_ptr(x2d.device)))
        return out.transpose(1, 2)

    def _conv_backward_raw(self, spec3d):
        n_sig, bins, frames = spec3d.shape
        out_len = max((frames + 1) * self.hop_length - self.frame_length, 0)
        if self.frame_length == self.hop_length:
            out_len = 0   # the reference slices ``x[..., 0:-0]`` = nothing (stft.py:296-298)
        y = torch.empty((n_sig, out_len), dtype=torch.float32, device=spec3d.device)
        if n_sig and out_len:
            with _lib.on_device(spec3d.device):
                _lib.check(_lib.lib().brv_convstft_backward(
                    self._core._plan(spec3d.device), _lib.ptr(spec3d), spec3d.stride(0),
                    spec3d.stride(1), spec3d.stride(2), n_sig, frames,
                    int(bool(self.normalized)), _lib.ptr(y), _lib.stream_ptr(spec3d.device)))
        return y

    def _gains(self):
        """(analysis gain, synthesis gain) the kernels apply besides scale_factor (stft.py:232-238,
        291-292): the two maps are adjoint up to these, which is what the gradients use."""
        nf = self._normalization_factor
        return (1.0 / nf, 1.0 / nf) if self.normalized else (1.0, 1.0 / (nf * nf))

    def backward(self, x, input_type='complex'):
        if input_type in ('real_imag', 'mag_phase'):
            a, b = x
            _lib.require_cuda(a, 'ConvSTFT.backward input')
            x = specfmt.join(a, b, input_type)
        elif input_type != 'complex':
            raise ValueError('input_type must be complex, real_imag or '
                             f'mag_phase, got {input_type}')
        _lib.require_cuda(x, 'ConvSTFT.backward input')
        if not x.is_complex():
            raise RuntimeError('ConvSTFT.backward input must be complex')
        lead = x.shape[:-2]
        spec3d = x.reshape(-1, *x.shape[-2:]).to(torch.complex64).resolve_conj().resolve_neg()
        if spec3d.shape[1] != self.frame_length // 2 + 1:
            raise RuntimeError(f'expected {self.frame_length // 2 + 1} frequency bins, got {spec3d.shape[1]}')
        if torch.is_grad_enabled() and spec3d.requires_grad:
            y = _ConvBackwardFunction.apply(spec3d, self)
        else:
            y = self._conv_backward_raw(spec3d)
        return y.view(*lead, -1)


def _conv_grad_check(conv):
    if conv.compression_factor != 1:
        raise NotImplementedError('gradient of the compressed ConvSTFT (compression_factor != 1) '
                                  'is not implemented')
    if conv.frame_length == conv.hop_length:
        raise NotImplementedError('ConvSTFT gradients need hop_length < frame_length (the synthesis '
                                  'returns an empty signal at hop_length == frame_length, stft.py:296-298)')


class _ConvForwardFunction(torch.autograd.Function):
    """X = s g_a A x; the synthesis kernel computes (g_s / s) Re(A^H X): dL/dx = s^2 g_a / g_s of it."""

    @staticmethod
    def forward(ctx, x2d, conv):
        _conv_grad_check(conv)
        ctx.conv, ctx.samples = conv, x2d.shape[-1]
        return conv._conv_forward_raw(x2d)

    @staticmethod
    def backward(ctx, grad):
        conv = ctx.conv
        ga, gs = conv._gains()
        g = grad.to(torch.complex64).resolve_conj().resolve_neg()
        L, H = conv.frame_length, conv.hop_length
        short = ctx.samples - ((g.shape[-1] + 1) * H - L)
        if short > 0:
            # hop does not divide 2 (L - H): the synthesis' trimmed output ends before the input did;
            # zero frames appended to the gradient extend it without changing the values
            g = torch.nn.functional.pad(g, (0, -(-short // H)))
        full = conv._conv_backward_raw(g)              # right-padded length >= samples
        return full[..., :ctx.samples] * (conv.scale_factor ** 2 * ga / gs), None


class _ConvBackwardFunction(torch.autograd.Function):
    """y = (g_s / s) Re(A^H X); the analysis kernel computes s g_a A y': dL/dX = g_s / (s^2 g_a) of it."""

    @staticmethod
    def forward(ctx, spec3d, conv):
        _conv_grad_check(conv)
        ctx.conv, ctx.frames, ctx.in_dtype = conv, spec3d.shape[-1], spec3d.dtype
        return conv._conv_backward_raw(spec3d)

    @staticmethod
    def backward(ctx, grad):
        conv = ctx.conv
        ga, gs = conv._gains()
        g = conv._conv_forward_raw(grad.float().contiguous())
        assert g.shape[-1] == ctx.frames
        return (g * (gs / (conv.scale_factor ** 2 * ga))).to(ctx.in_dtype), None


class _MelApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fb, inverse):
        ctx.fb, ctx.inverse = fb, inverse
        return fb._apply_raw(x, 'inverse' if inverse else 'forward')

    @staticmethod
    def backward(ctx, grad):
        which = 'inverse_t' if ctx.inverse else 'forward_t'
        return ctx.fb._apply_raw(grad, which), None, None


class MelFilterbank:
    """Area-normalised triangular HTK-mel filterbank (stft.py:152-198).

    ``filters`` / ``fc`` / ``scaling`` / ``inverse_filters`` are the same float32
    tensors the reference builds; ``forward`` / ``backward`` apply them as banded
    (CSR) gathers on the GPU — each bin feeds at most two filters, so a dense
    matmul over the 97 %-zero matrix is never formed.
    """

    def __init__(self, n_filters=64, n_fft=512, fs=16e3, fmin=50, fmax=8000):
        self.n_filters = n_filters
        self.n_fft = n_fft
        self.fs = fs
        self.fmin = fmin
        self.fmax = fmax
        self.filters, self.fc, self.scaling = self.calc_filterbank()
        self._csr = {}

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_csr'] = {}
        return state

    def calc_filterbank(self):
        # Same elementary float32 torch operations as stft.py:161-176 (so the
        # constants are bit-identical), evaluated for all filters at once.
        mel = torch.linspace(self.freq_to_mel(self.fmin),
                             self.freq_to_mel(self.fmax), self.n_filters + 2)
        fc = self.mel_to_freq(mel)
        f = torch.from_numpy(_fft_freqs(self.fs, self.n_fft)).float()
        lo, mid, hi = fc[:-2, None], fc[1:-1, None], fc[2:, None]
        rising = (f[None] - lo) / (mid - lo)
        falling = (hi - f[None]) / (hi - mid)
        filters = torch.zeros((self.n_filters, len(f)))
        filters = torch.where((lo <= f[None]) & (f[None] <= mid), rising, filters)
        filters = torch.where((mid <= f[None]) & (f[None] <= hi), falling, filters)
        scaling = filters.sum(axis=1, keepdims=True)
        filters = filters / scaling
        return filters, fc, scaling

    @staticmethod
    def mel_to_freq(mel):
        return 700 * (10 ** (mel / 2595) - 1)

    @staticmethod
    def freq_to_mel(f):
        return 2595 * math.log10(1 + f / 700)

    @property
    def inverse_filters(self):
        return (self.filters * self.scaling).T

    def csr(self, which, device):
        """(vals, cols, rowptr, n_rows_out, n_rows_in) of a dense matrix, on device."""
        key = (which, device.index if device.index is not None
               else torch.cuda.current_device())
        if key not in self._csr:
            dense = {'forward': self.filters,
                     'inverse': self.inverse_filters,
                     'forward_t': self.filters.T,
                     'inverse_t': self.inverse_filters.T}[which]
            dense = dense.contiguous()
            nz = dense != 0
            counts = nz.sum(1)
            rowptr = torch.zeros(dense.shape[0] + 1, dtype=torch.int32)
            rowptr[1:] = counts.cumsum(0)
            rows, cols = nz.nonzero(as_tuple=True)
            self._csr[key] = (dense[rows, cols].to(device),
                              cols.to(torch.int32).to(device),
                              rowptr.to(device), dense.shape[0], dense.shape[1])
        return self._csr[key]

    def _apply_raw(self, x, which):
        _lib.require_cuda(x, 'MelFilterbank input')
        if x.dtype != torch.float32:
            x = x.float()
        vals, cols, rowptr, n_out, n_in = self.csr(which, x.device)
        if x.ndim < 2 or x.shape[-2] != n_in:
            raise RuntimeError(f'expected {n_in} rows on dim -2, got {tuple(x.shape)}')
        lead = x.shape[:-2]
        frames = x.shape[-1]
        x3 = x.reshape(-1, n_in, frames)
        out = torch.empty((x3.shape[0], n_out, frames), dtype=torch.float32,
                          device=x.device)
        # grid.z carries the batch: chunk to the CUDA limit
        for start in range(0, x3.shape[0], 65535):
            chunk = x3[start:start + 65535]
            with _lib.on_device(x.device):
                _lib.check(_lib.lib().brv_mel_apply(
                    _lib.ptr(chunk), chunk.stride(0), chunk.stride(1),
                    chunk.stride(2), chunk.shape[0], n_in, frames,
                    _lib.ptr(vals), _lib.ptr(cols), _lib.ptr(rowptr), n_out,
                    _lib.ptr(out[start:start + 65535]),
                    _lib.stream_ptr(x.device)))
        return out.view(*lead, n_out, frames)

    def __call__(self, x):
        return self.forward(x)

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            return _MelApply.apply(x, self, False)
        return self._apply_raw(x, 'forward')

    def backward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            return _MelApply.apply(x, self, True)
        return self._apply_raw(x, 'inverse')
