"""MANNER's multi-resolution STFT loss on the sm_100a kernels.

Mirror of ``brever/models/manner/stft_loss.py:22-151`` (the Demucs / ParallelWaveGAN loss):
per resolution ``torch.stft(x, fft_size, hop_size, win_length, hann_window)`` -- centre
reflect padding, raw torch framing (no ``STFT.pad``), no normalisation --, clamp-sqrt
magnitudes, spectral convergence ``||Y| - |X||_F / ||Y||_F`` and log-magnitude L1, averaged
over the resolutions and scaled by ``factor_sc`` / ``factor_mag``.  Same class names,
constructor arguments and return values (two ``(B,)`` tensors).

Data flow per resolution: one reflect-pad gather kernel per signal batch, the STFT kernels
(folded tcgen05 for ``n_fft <= 512``, the dense tcgen05 contraction for 1024 / 2048), one
reduction kernel that reads both spectrograms once and produces the three sums the two
losses need (no magnitude, log or difference tensor is materialised), and for the gradient
one elementwise kernel into the STFT-gradient kernel.
"""
import torch

from . import _lib
from .modules import STFT


class _SpectralLossFunction(torch.autograd.Function):
    """(X, Y) complex64 (B, F, T) views of frame-major memory -> sc (B,), mag (B,)."""

    @staticmethod
    def forward(ctx, X, Y):
        batch = X.shape[0]
        n_elems = X[0].numel() if batch else 0
        sums = torch.empty((batch, 3), dtype=torch.float64, device=X.device)
        if batch:
            with _lib.on_device(X.device):
                _lib.check(_lib.lib().brv_mrstft_forward(
                    _lib.ptr(X), _lib.ptr(Y), batch, n_elems, _lib.ptr(sums),
                    _lib.stream_ptr(X.device)))
        sc = (sums[:, 0] / sums[:, 1]).sqrt()
        mag = sums[:, 2] / max(n_elems, 1)
        ctx.save_for_backward(X, Y, sums)
        ctx.n_elems = n_elems
        return sc.float(), mag.float()

    @staticmethod
    def backward(ctx, g_sc, g_mag):
        X, Y, sums = ctx.saved_tensors
        batch = X.shape[0]
        k_sc = (g_sc.double() / (sums[:, 0] * sums[:, 1]).sqrt()).float().contiguous()
        k_sc = torch.nan_to_num(k_sc, nan=0.0, posinf=0.0, neginf=0.0)     # identical spectrograms
        k_mag = (g_mag.float() / max(ctx.n_elems, 1)).contiguous()
        gX = torch.empty_strided(X.shape, X.stride(), dtype=torch.complex64, device=X.device)
        if batch:
            with _lib.on_device(X.device):
                _lib.check(_lib.lib().brv_mrstft_backward(
                    _lib.ptr(X), _lib.ptr(Y), _lib.ptr(k_sc), _lib.ptr(k_mag), batch, ctx.n_elems,
                    _lib.ptr(gX), _lib.stream_ptr(X.device)))
        return gX, None


def _dense_spec(spec):
    """The (B, T, F) memory behind STFT.forward's (B, F, T) view, as a dense tensor."""
    mem = spec.transpose(-1, -2)
    return mem if mem.is_contiguous() else mem.contiguous()


class STFTLoss:
    """One resolution (stft_loss.py:80-106)."""

    def __init__(self, fft_size=1024, shift_size=120, win_length=600, window='hann_window'):
        self.fft_size, self.shift_size, self.win_length = fft_size, shift_size, win_length
        self.window = getattr(torch, window)(win_length)          # periodic, as torch.stft's callers build it
        self.stft = STFT(frame_length=win_length, hop_length=shift_size, n_fft=fft_size,
                         window=self.window, normalized=False, pad_mode='reflect')
        self.stft._raw_framing = True                             # torch.stft, not brever's STFT.pad

    def forward(self, x, y):
        X = _dense_spec(self.stft(x))
        Y = _dense_spec(self.stft(y))
        return _SpectralLossFunction.apply(X, Y.detach())

    __call__ = forward


class MultiResolutionSTFTLoss:
    """stft_loss.py:109-151: mean over resolutions of (spectral convergence, log-magnitude)."""

    def __init__(self, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50],
                 win_lengths=[600, 1200, 240], window='hann_window', factor_sc=0.1,
                 factor_mag=0.1):
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = [STFTLoss(fs, ss, wl, window)
                            for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths)]
        self.factor_sc = factor_sc
        self.factor_mag = factor_mag

    def forward(self, x, y):
        sc_loss, mag_loss = 0.0, 0.0
        for f in self.stft_losses:
            sc_l, mag_l = f(x, y)
            sc_loss = sc_loss + sc_l
            mag_loss = mag_loss + mag_l
        sc_loss = sc_loss / len(self.stft_losses)
        mag_loss = mag_loss / len(self.stft_losses)
        return self.factor_sc * sc_loss, self.factor_mag * mag_loss

    __call__ = forward
