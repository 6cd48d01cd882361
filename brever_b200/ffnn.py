"""FFNN feature glue (brever/models/ffnn/ffnn.py:77-203) on the sm_100a kernels.

The reference keeps these as methods of the ``FFNN`` model; the network itself
(``_FFNN``, ffnn.py:151-172) is out of scope and stays in PyTorch, so this module
exposes the glue as free functions / small modules with the same names and
semantics, plus ``FFNNFrontEnd`` which chains STFT -> log-mel -> context stack ->
static normalisation (and the ``_enhance`` tail: mask extrapolation -> iSTFT)
through the fused kernels.  A maintainer swaps the bodies of ``FFNN.stack``,
``FFNN.irm`` ... for these calls.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .modules import STFT, FeatureExtractor, MelFilterbank

eps = np.finfo(float).eps  # ffnn.py:12 (float64 eps, 2.2e-16)


def _stats(t, rows, device):
    if t is None:
        return None
    t = t.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    if t.numel() != rows:
        raise RuntimeError(f'statistics must have {rows} entries, got {t.numel()}')
    return t


def _stack_raw(data, stacks, decimation=1, mean=None, std=None):
    _lib.require_cuda(data, 'stack input')
    unbatched = data.ndim == 2
    x = data.unsqueeze(0) if unbatched else data
    if x.ndim != 3:
        raise ValueError(f'input must be 2 or 3 dimensional, got {data.ndim}')
    x = x.float().contiguous()
    batch, nf, frames = x.shape
    rows = nf * (stacks + 1)
    out_frames = -(-frames // decimation)
    out = torch.empty((batch, rows, out_frames), dtype=torch.float32, device=x.device)
    mean_t, std_t = _stats(mean, rows, x.device), _stats(std, rows, x.device)
    for start in range(0, batch, 65535):
        chunk = x[start:start + 65535]
        with _lib.on_device(x.device):
            _lib.check(_lib.lib().brv_stack_normalize(
                _lib.ptr(chunk), chunk.shape[0], nf, frames, int(stacks),
                int(decimation), _lib.ptr(mean_t), _lib.ptr(std_t),
                _lib.ptr(out[start:start + 65535]), _lib.stream_ptr(x.device)))
    return out[0] if unbatched else out


def stack(data, stacks):
    """``out[k*nf + f, t] = data[f, max(t - k, 0)]``, k = 0..stacks (ffnn.py:122-132)."""
    return _stack_raw(data, stacks)


def decimate(data, decimation):
    """ffnn.py:134-135."""
    return data[..., ::decimation]


def irm(foreground_mag, background_mag, mel_fb):
    """Ideal ratio mask labels on ``(C, F, T)`` magnitudes (ffnn.py:113-120)."""
    fg = mel_fb(foreground_mag.pow(2).mean(0))
    bg = mel_fb(background_mag.pow(2).mean(0))
    return (1 + bg / (fg + eps)).pow(-0.5)


def channel_mean(spec, mask=None):
    """``spec.mean(1) [* mask]`` of a ``(B, C, F, T)`` complex STFT with a real ``(B, F, T)`` mask,
    in one pass (``FFNN._enhance``, ffnn.py:107-110).  Returns the ``(B, F, T)`` view of
    frame-major memory that ``STFT.backward`` streams.  Not differentiable: inference path."""
    _lib.require_cuda(spec, 'channel_mean input')
    if spec.ndim != 4 or not spec.is_complex():
        raise ValueError(f'expected a complex (B, C, F, T) tensor, got {tuple(spec.shape)}')
    spec = spec.to(torch.complex64).resolve_conj().resolve_neg()
    batch, channels, bins, frames = spec.shape
    ms = (0, 0, 0)
    if mask is not None:
        _lib.require_cuda(mask, 'channel_mean mask')
        mask = mask.float().expand(batch, bins, frames)
        ms = mask.stride()
    out = torch.empty((batch, frames, bins), dtype=torch.complex64, device=spec.device)
    for start in range(0, batch, 65535):
        sub, msub = spec[start:start + 65535], None if mask is None else mask[start:start + 65535]
        with _lib.on_device(spec.device):
            _lib.check(_lib.lib().brv_channel_mean_mask(
                _lib.ptr(sub), *sub.stride(), _lib.ptr(msub), *ms, sub.shape[0], channels, bins, frames,
                _lib.ptr(out[start:start + 65535]), _lib.stream_ptr(spec.device)))
    return out.transpose(1, 2)


def accumulate_mean(total, values):
    """``total += values.mean()`` in one launch (the trainer's running loss, training.py:369-373)."""
    v = values.detach().float().contiguous()
    with _lib.on_device(v.device):
        _lib.check(_lib.lib().brv_accumulate_mean(_lib.ptr(v), v.numel(), _lib.ptr(total),
                                                  _lib.stream_ptr(v.device)))
    return total


def _frame_mask(x, frames):
    """Zero the columns t >= frames[b] of ``(B, ..., T)`` (what collating per-item outputs does)."""
    flat = x.reshape(x.shape[0], -1, x.shape[-1]).contiguous()
    out = torch.empty_like(flat)
    if flat.numel():
        with _lib.on_device(flat.device):
            _lib.check(_lib.lib().brv_apply_mask(
                _lib.ptr(flat), _lib.ptr(frames), flat.shape[0], flat.shape[1], flat.shape[2],
                _lib.ptr(out), _lib.stream_ptr(flat.device)))
    return out.view(x.shape)


class StaticNormalizer(nn.Module):
    """``(x - mean) / std`` with ``(input_size, 1)`` buffers (ffnn.py:175-187)."""

    def __init__(self, input_size):
        super().__init__()
        self.register_buffer('mean', torch.zeros((input_size, 1)))
        self.register_buffer('std', torch.ones((input_size, 1)))

    def set_statistics(self, mean, std):
        self.mean[:], self.std[:] = mean, std

    def forward(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            return (x - self.mean) / self.std   # autograd path: plain broadcast
        return _stack_raw(x, 0, 1, self.mean, self.std)


class CumulativeNormalizer(nn.Module):
    """Running mean / variance normalisation along frames (ffnn.py:190-203)."""

    def __init__(self, eps=1e-4):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        _lib.require_cuda(x, 'CumulativeNormalizer input')
        src = x.float().contiguous()
        out = torch.empty_like(src)
        if src.numel():
            with _lib.on_device(src.device):
                _lib.check(_lib.lib().brv_cumulative_normalize(
                    _lib.ptr(src), src.numel() // src.shape[-1], src.shape[-1],
                    float(self.eps), _lib.ptr(out), _lib.stream_ptr(src.device)))
        return out


def training_statistics(items):
    """FFNN.pre_train statistics (ffnn.py:137-148): unweighted mean over items of
    the per-item time mean and mean square -> ``(mean, std)``."""
    mean, msq = 0, 0
    for data in items:
        mean = mean + data.mean(-1, keepdim=True)
        msq = msq + data.pow(2).mean(-1, keepdim=True)
    mean, msq = mean / len(items), msq / len(items)
    return mean, (msq - mean.pow(2)).sqrt()


class FFNNFrontEnd:
    """STFT -> log-mel -> stack -> (static) normalise, and the enhance tail.

    Mirrors the data flow of ``FFNN.transform`` / ``FFNN._enhance``
    (ffnn.py:77-111) with the kernels fused: the feature kernel reads the
    spectrogram once and writes the stacked, normalised features directly.
    """

    def __init__(self, fs=16000, features=('logfbe',), stacks=5, decimation=1,
                 stft_frame_length=512, stft_hop_length=256, stft_window='hann',
                 mel_filters=64):
        self.stacks = stacks
        self.decimation = decimation
        self.stft = STFT(frame_length=stft_frame_length,
                         hop_length=stft_hop_length, window=stft_window)
        self.mel_fb = MelFilterbank(n_filters=mel_filters,
                                    n_fft=stft_frame_length, fs=fs)
        self.feature_extractor = FeatureExtractor(
            features=set(features), mel_fb=self.mel_fb,
            hop_length=stft_hop_length, fs=fs)
        self.input_size = self.feature_extractor.n_features * (stacks + 1)

    def features(self, spec, mean=None, std=None):
        """(B, C, F, T) complex -> (B, input_size, T') stacked [normalised] features."""
        names = self.feature_extractor.features
        if len(names) == 1 and names[0] in FeatureExtractor._FBE_FAMILY:
            normalize, compression = FeatureExtractor._FBE_FAMILY[names[0]]
            return self.feature_extractor.fbe(
                spec, normalize=normalize, compression=compression,
                stacks=self.stacks, decimation=self.decimation, mean=mean,
                std=std)
        feats = torch.cat([self.feature_extractor.calc_feature(spec, n)
                           for n in names], dim=1)
        return _stack_raw(feats, self.stacks, self.decimation, mean, std)

    def transform(self, sources):
        """FFNN.transform (ffnn.py:77-91) for one ``(2, C, L)`` utterance."""
        assert sources.shape[0] == 2  # mixture, foreground
        spec = self.stft(sources)
        mix, foreground = spec
        background = mix - foreground
        x = self.features(mix.unsqueeze(0))[0]
        labels = irm(foreground.abs(), background.abs(), self.mel_fb)
        labels = decimate(labels, self.decimation)
        return torch.cat([x, labels])

    def transform_batched(self, batch, lengths):
        """``FFNN.transform`` for a collated device batch, in one pass of the kernels.

        The reference transforms a validation batch one utterance at a time on the device
        and re-collates the results (``training.py:336-338``, ``data.py:408-491``); here
        ``batch`` is the zero-padded ``(B, 2, C, L)`` tensor of (mixture, foreground) pairs
        with ``lengths`` valid samples per item, and the result is what that loop followed
        by ``_collate_fn`` returns: ``(B, input_size + n_labels, T')`` with the columns past
        each item's own frame count zeroed, and the per-item frame counts."""
        if batch.ndim != 4 or batch.shape[1] != 2:
            raise ValueError(f'batch must be (B, 2, channels, samples), got {tuple(batch.shape)}')
        lengths = torch.as_tensor(lengths).to(device=batch.device, dtype=torch.int64)
        spec = self.stft(batch)                                  # (B, 2, C, F, T)
        mix, foreground = spec[:, 0], spec[:, 1]
        background = mix - foreground
        x = self.features(mix)                                   # (B, input_size, T')
        fe = self.feature_extractor
        fg = fe.fbe(foreground, normalize=False, compression='none')     # mel(mean_c |X|^2)
        bg = fe.fbe(background, normalize=False, compression='none')
        labels = decimate((1 + bg / (fg + eps)).pow(-0.5), self.decimation)
        out = torch.cat([x, labels], dim=1)
        hop, fl = self.stft.hop_length, self.stft.frame_length
        frames0 = (torch.clamp(lengths - fl, min=0) + hop - 1) // hop + 1       # STFT.frame_count
        frames = 1 + ((frames0 - 1) * hop + fl + 2 * (self.stft.n_fft // 2) - self.stft.n_fft) // hop
        out_frames = (frames + self.decimation - 1) // self.decimation
        return _frame_mask(out, out_frames), out_frames

    def enhance(self, x, mask_fn, mean=None, std=None):
        """FFNN._enhance (ffnn.py:100-111); ``mask_fn`` is the network."""
        length = x.shape[-1]
        spec = self.stft(x)
        feats = self.features(spec, mean, std)
        mask = mask_fn(feats)
        extrapolated = self.mel_fb.backward(mask)
        out = self.stft.backward(channel_mean(spec, extrapolated))     # one pass: mean_c X * mask
        return out[..., :length]
