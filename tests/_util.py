"""Shared helpers for the tests: seeded inputs (same recipe as
tests/golden/make_golden.py) and the parity metric of SURVEY §8(d)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                      'reference_vectors.npz')
_cache = {}


def golden():
    if 'g' not in _cache:
        _cache['g'] = np.load(GOLDEN)
    return _cache['g']


def randn(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


def crandn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    re = torch.randn(*shape, generator=g)
    im = torch.randn(*shape, generator=g)
    return torch.complex(re, im)


def rel_err(new, ref):
    """(max-norm error / max|ref|, Frobenius error / ||ref||_F)."""
    new = np.asarray(new)
    ref = np.asarray(ref)
    assert new.shape == ref.shape, (new.shape, ref.shape)
    if ref.size == 0:
        return 0.0, 0.0
    diff = np.abs(new.astype(np.complex128) - ref.astype(np.complex128))
    scale = max(np.abs(ref).max(), 1e-30)
    fro = np.sqrt((diff ** 2).sum()) / max(np.sqrt((np.abs(ref) ** 2).sum()),
                                            1e-30)
    return float(diff.max() / scale), float(fro)


def assert_parity(new, ref, tol=1e-4, what=''):
    """north_star tolerance: 1e-4 relative (fp32), shapes exact."""
    mx, fro = rel_err(new, ref)
    assert mx <= tol and fro <= tol, f'{what}: max-rel {mx:.3e}, fro {fro:.3e}'


STFT_SHAPE_CASES = [
    (100, 512, 256, None, 'hann'), (4000, 512, 256, None, 'hann'),
    (777, 512, 128, None, 'hann'), (1000, 256, 128, None, 'hann'),
    (3001, 510, 128, None, 'hann'), (1600, 400, 100, 512, 'hann'),
    (2048, 512, 256, None, 'hamming'), (900, 256, 64, None, None),
    (512, 512, 256, None, 'hann'), (513, 512, 256, None, 'hann'),
]


def synthetic_mixture(shape, seed, fs=16000):
    """SURVEY §8(d) synthetic mixtures: low-passed 0.05*randn foreground plus
    white noise at an SNR drawn from U(-5, 10) dB.  CPU generator, float32."""
    g = torch.Generator().manual_seed(seed)
    fg = 0.05 * torch.randn(*shape, generator=g)
    # one-pole low-pass (speech-like colouring), done as a cumulative filter
    # in the frequency domain to stay vectorised
    n = shape[-1]
    spec = torch.fft.rfft(fg, dim=-1)
    w = torch.arange(spec.shape[-1]) * (2 * np.pi / n)
    pole = 0.9
    h = (1 - pole) / (1 - pole * torch.exp(-1j * w))
    fg = torch.fft.irfft(spec * h, n=n, dim=-1) * 4.0
    snr_db = torch.rand(shape[:-1] + (1,), generator=g) * 15 - 5
    noise = 0.05 * 10 ** (-snr_db / 20) * torch.randn(*shape, generator=g)
    return (fg + noise).float(), fg.float()
