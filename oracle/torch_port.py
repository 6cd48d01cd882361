"""PyTorch-CPU port of the reference hot path (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference (pure Python) runs this path through ``torch.stft`` / ``torch.istft``
/ ``torch.matmul`` on whatever device its tensors live on; on a host without
brever installed (the GPU box has no ``/root/reference``) this file restates
those exact library calls, in float32 on CPU, so that ``bench.py`` can time "the
reference's CPU implementation" on the box's own cores (``cpu_baseline.kind =
"port"``) and tests can compare like with like.  Nothing under ``brever_b200/``
imports it.

Pinned against the real reference by ``tests/test_oracle_golden.py`` (golden
vectors produced by ``tests/golden/make_golden.py`` from ``/root/reference``).

Each function cites the reference lines it follows.
"""
import math
from itertools import permutations

import torch
import torch.nn.functional as F

EPS32 = torch.finfo(torch.float32).eps


def _frames0(samples, frame_length, hop_length):
    # stft.py:146-149
    return math.ceil(max(samples - frame_length, 0) / hop_length) + 1


def stft(x, window, frame_length=512, hop_length=256, normalized=True,
         onesided=True, compression_factor=1, scale_factor=1, n_fft=None):
    """stft.py:59-89.  ``window``: float64 tensor of ``frame_length`` points."""
    n_fft = frame_length if n_fft is None else n_fft
    win = window.to(dtype=x.dtype, device=x.device)
    samples = x.shape[-1]
    extra = (_frames0(samples, frame_length, hop_length) - 1) * hop_length \
        + frame_length - samples
    x = F.pad(x, (0, extra))
    lead = x.shape[:-1]
    spec = torch.stft(x.reshape(-1, x.shape[-1]), n_fft=n_fft,
                      hop_length=hop_length, win_length=frame_length,
                      window=win, center=True, pad_mode='constant',
                      normalized=False, onesided=onesided,
                      return_complex=True)
    if normalized:
        spec /= win.pow(2).sum().sqrt()
    if compression_factor != 1:
        spec = spec.abs().pow(compression_factor) \
            * torch.exp(1j * spec.angle())
    spec *= scale_factor
    return spec.view(*lead, *spec.shape[-2:])


def istft(spec, window, frame_length=512, hop_length=256, normalized=True,
          onesided=True, compression_factor=1, scale_factor=1, n_fft=None):
    """stft.py:111-138 (out of place: the caller's tensor is left untouched)."""
    n_fft = frame_length if n_fft is None else n_fft
    win = window.to(dtype=spec.real.dtype, device=spec.device)
    spec = spec / scale_factor
    if compression_factor != 1:
        spec = spec.abs().pow(1 / compression_factor) \
            * torch.exp(1j * spec.angle())
    if normalized:
        spec = spec * win.pow(2).sum().sqrt()
    lead = spec.shape[:-2]
    out = torch.istft(spec.reshape(-1, *spec.shape[-2:]), n_fft=n_fft,
                      hop_length=hop_length, win_length=frame_length,
                      window=win, center=True, normalized=False,
                      onesided=onesided, return_complex=False)
    return out.view(*lead, -1)


def logfbe(spec, filters, normalize=False, compression='log'):
    """features.py:186-198 on a (B, C, F, T) complex spectrogram."""
    out = spec.abs().pow(2).mean(1)
    out = torch.matmul(filters.to(out.device), out)
    if normalize:
        out /= out.sum(1, keepdims=True) + EPS32
    if compression == 'log':
        out = torch.log(out + EPS32)
    elif compression == 'cubic':
        out = out.pow(1 / 3)
    return out


def stack(data, stacks):
    """ffnn.py:122-132."""
    pieces = [data]
    for k in range(1, stacks + 1):
        shifted = data.roll(k, -1)
        shifted[..., :k] = data[..., :1]
        pieces.append(shifted)
    return torch.cat(pieces, dim=0 if data.ndim == 2 else 1)


def static_normalize(x, mean, std):
    """ffnn.py:186-187."""
    return (x - mean) / std


def mask_pair(x, y, lengths):
    """criterion.py:229-234."""
    mask = torch.zeros(x.shape, device=x.device)
    for i, n in enumerate(lengths):
        mask[i, ..., :n] = 1
    return x * mask, y * mask


def snr(x, y, lengths):
    """criterion.py:96-101."""
    x, y = mask_pair(x, y, lengths)
    ratio = y.pow(2).sum(-1) / ((y - x).pow(2).sum(-1) + EPS32)
    db = 10 * torch.log10(ratio + EPS32)
    return -db.mean(tuple(range(1, x.ndim - 1)))


def sisnr(x, y, lengths):
    """criterion.py:41-72 (out-of-place division at the end, see SURVEY §8a')."""
    x, y = mask_pair(x, y, lengths)
    x = x - x.sum(2, keepdim=True) / lengths.view(-1, 1, 1)
    y = y - y.sum(2, keepdim=True) / lengths.view(-1, 1, 1)
    x, y = mask_pair(x, y, lengths)
    est = x.unsqueeze(1)
    ref = y.unsqueeze(2)
    proj = (est * ref).sum(3, keepdim=True) * ref \
        / ref.pow(2).sum(3, keepdim=True)
    noise = est - proj
    db = 10 * torch.log10(proj.pow(2).sum(3)
                          / (noise.pow(2).sum(3) + EPS32) + EPS32)
    n_src = x.shape[1]
    perms = torch.tensor(list(permutations(range(n_src))), dtype=torch.long)
    one_hot = x.new_zeros((*perms.shape, n_src)).scatter_(
        2, perms.unsqueeze(2), 1)
    best = torch.einsum('bij,pij->bp', db, one_hot).amax(1)
    return -(best / n_src)


def mse(x, y, lengths, weight=None):
    """criterion.py:125-132 (out of place)."""
    x, y = mask_pair(x, y, lengths)
    loss = (x - y).abs().pow(2).sum(-1)
    loss = loss / lengths.view(-1, *[1] * (x.ndim - 2))
    if weight is not None:
        loss = loss * weight.view(-1, *[1] * (x.ndim - 2))
    return loss.mean(tuple(range(1, x.ndim - 1)))


def multiresyu(x, y, lengths, frame_lengths=(512,), hop_lengths=None,
               time_domain_weight=0.5, spectral_weight=0.5, scale_invariant=False):
    """criterion.py:164-226: boxcar, unnormalised STFT magnitudes + L1 in time."""
    if hop_lengths is None:
        hop_lengths = [n // 2 for n in frame_lengths]
    x, y = mask_pair(x, y, lengths)
    if scale_invariant:
        scaling = (x * y).sum(-1, keepdim=True) / (x.pow(2).sum(-1, keepdim=True) + EPS32)
    else:
        scaling = 1
    out = time_domain_weight * (scaling * x - y).abs().sum(-1)
    for n, hop in zip(frame_lengths, hop_lengths):
        win = torch.ones(n, dtype=torch.float64)
        kw = dict(frame_length=n, hop_length=hop, normalized=False)
        y_mag = stft(y, win, **kw).abs()
        x_mag = stft(scaling * x, win, **kw).abs()
        out = out + spectral_weight * (x_mag - y_mag).abs().sum((-2, -1)) / len(frame_lengths)
    out = out / lengths.view(-1, *[1] * (x.ndim - 2))
    return out.mean(tuple(range(1, x.ndim - 1)))
