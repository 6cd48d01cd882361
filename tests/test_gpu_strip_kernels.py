"""GPU parity tests for the transposed strip kernels (brv_fold_t.cuh): the forward kernel the
default dispatch picks for large launches (variant 5 forces it at any size) and the inverse kernel
it picks for n_fft = 512-class geometries (variants 6 / 7 force it with 64- / 32-frame tiles at any
size and geometry), against the float64 oracle, the one-tile-per-TMEM kernels (variant 4) and
the generic path -- forward, inverse (both layouts), both gradients, ConvSTFT, ragged strip ends."""
import numpy as np
import pytest
import torch

import brever_b200 as brv
from brever_b200 import _lib
from oracle import tf_oracle as O

from _util import crandn, randn, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


class variant:
    def __init__(self, v):
        self.v = v

    def __enter__(self):
        self.prev = _lib.lib().brv_set_tc_variant(self.v)

    def __exit__(self, *exc):
        _lib.lib().brv_set_tc_variant(self.prev)


def cpu(t):
    return t.detach().cpu().numpy()


FWD_CASES = [
    dict(frame_length=512, hop_length=128),
    dict(frame_length=512, hop_length=256),
    dict(frame_length=256, hop_length=128, normalized=False),
    dict(frame_length=128, hop_length=32),
    dict(frame_length=384, hop_length=96, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=400, hop_length=100, n_fft=512),
    dict(frame_length=512, hop_length=100, window='hamming'),
    dict(frame_length=512, hop_length=512),
]


@pytest.mark.parametrize('kw', FWD_CASES)
@pytest.mark.parametrize('shape', [(1, 100), (3, 4097), (37, 64000), (400, 20000)])
def test_strip_forward(kw, shape):
    x = randn(shape, 5)
    x[::3] *= 1e-3
    x[1::3] *= 200.0
    stft = brv.STFT(**kw)
    with variant(5):
        new = stft(x.to(DEV))
        again = stft(x.to(DEV))
    with variant(4):
        old = stft(x.to(DEV))
    assert new.shape == old.shape and new.stride() == old.stride()
    assert torch.equal(new, again)                     # deterministic
    for i in range(0, shape[0], max(1, shape[0] // 7)):
        e = rel_err(cpu(new[i]), cpu(old[i]))
        assert e[0] < 2e-6, (kw, shape, i, e)
        e = rel_err(cpu(new[i]), O.stft(x[i].numpy(), **kw))
        assert e[0] < 1e-4 and e[1] < 1e-4, (kw, shape, i, e)


def test_strip_forward_nan_stays_in_its_frames():
    stft = brv.STFT(512, 128)
    x = randn((2, 20000), 3)
    x[1, 7000] = float('nan')
    with variant(5):
        spec = stft(x.to(DEV))
    bad = torch.isnan(spec[1].real).any(0).cpu().numpy()
    frames = np.nonzero(bad)[0]
    assert frames.min() >= (7000 + 256 - 511) // 128 and frames.max() <= (7000 + 256) // 128
    assert not torch.isnan(spec[0].real).any()


INV_CASES = [
    dict(frame_length=512, hop_length=128),
    dict(frame_length=512, hop_length=256),
    dict(frame_length=256, hop_length=128, normalized=False),
    dict(frame_length=256, hop_length=64, window='hamming'),
    dict(frame_length=128, hop_length=32),
    dict(frame_length=384, hop_length=96, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=400, hop_length=128, n_fft=512),
    dict(frame_length=510, hop_length=128, normalized=False),                # n_fft = 4Q - 2
    dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=254, hop_length=64),
]


@pytest.mark.parametrize('kw', INV_CASES)
@pytest.mark.parametrize('shape', [(1, 2), (3, 9), (2, 130), (5, 501), (170, 67), (64, 200)])
@pytest.mark.parametrize('layout', ['bin_major', 'frame_major'])
@pytest.mark.parametrize('vi', [6, 7])
def test_strip_inverse(kw, shape, layout, vi):
    n_sig, frames = shape
    stft = brv.STFT(**kw)
    spec = crandn((n_sig, stft.n_bins, frames), 91)
    spec[::3] *= 1e-3
    spec[1::3] *= 300.0
    dev = spec.to(DEV)
    if layout == 'frame_major':
        dev = dev.transpose(1, 2).contiguous().transpose(1, 2)
    try:
        ref0 = O.istft(spec[0].numpy(), **kw)
    except RuntimeError:
        pytest.skip('NOLA')
    with variant(vi):
        new = stft.backward(dev)
        again = stft.backward(dev)
    with variant(4):
        old = stft.backward(dev)
    assert new.shape == old.shape
    assert torch.equal(new, again)
    for i in range(0, n_sig, max(1, n_sig // 6)):
        ref = ref0 if i == 0 else O.istft(spec[i].numpy(), **kw)
        e = rel_err(cpu(new[i]), ref)
        assert e[0] < 1e-4 and e[1] < 1e-4, (kw, shape, layout, i, e)
        e = rel_err(cpu(new[i]), cpu(old[i]))
        assert e[0] < 5e-6, (kw, shape, layout, i, e)


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128),
                                dict(frame_length=512, hop_length=256, scale_factor=0.3),
                                dict(frame_length=256, hop_length=128, normalized=False),
                                dict(frame_length=510, hop_length=128, normalized=False)])
@pytest.mark.parametrize('shape', [(3, 130), (160, 157)])
def test_strip_gradients(kw, shape):
    """d iSTFT / dX on the forward strip kernel (envelope summed on the fly by its loader warp)
    and d STFT / dx on the inverse strip kernel, against the one-tile-per-TMEM kernels."""
    n_sig, frames = shape
    stft = brv.STFT(**kw)
    spec = crandn((n_sig, stft.n_bins, frames), 31)
    v = randn((n_sig, stft.hop_length * (frames - 1)), 32)

    def grads(vf, vi):
        sg = spec.clone().to(DEV).requires_grad_(True)
        with variant(vi):
            y = stft.backward(sg)
        with variant(vf):
            (y * v.to(DEV)).sum().backward()
        xg = v.clone().to(DEV).requires_grad_(True)
        with variant(vf):
            X = stft(xg)
        with variant(vi):
            (X.real * spec.real.to(DEV) + X.imag * spec.imag.to(DEV)).sum().backward()
        return sg.grad, xg.grad

    g_old = grads(4, 4)
    for vi in (6, 7):
        g_new = grads(5, vi)
        for a, b, what in zip(g_new, g_old, ('d istft / dX', 'd stft / dx')):
            for i in range(0, n_sig, max(1, n_sig // 5)):
                e = rel_err(cpu(a[i]), cpu(b[i]))
                assert e[0] < 2e-5 and e[1] < 2e-5, (what, vi, kw, shape, i, e)


def test_strip_conv_stft_round_trip():
    conv = brv.ConvSTFT(frame_length=512, hop_length=128)
    x = randn((5, 30000), 7).to(DEV)
    with variant(4):
        X0 = conv(x)
        y0 = conv.backward(X0)
    with variant(5):
        X1 = conv(x)
    with variant(6):
        y1 = conv.backward(X0)
    assert rel_err(cpu(X1), cpu(X0))[0] < 2e-6
    assert rel_err(cpu(y1), cpu(y0))[0] < 5e-6


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128),
                                dict(frame_length=256, hop_length=128, normalized=False),
                                dict(frame_length=510, hop_length=128, normalized=False)])
@pytest.mark.parametrize('shape', [(1, 2_000_000), (4096, 1500), (7, 123_457)])
def test_round_trip_extreme_aspect_ratios(kw, shape):
    """Default dispatch, size-independent property: iSTFT(STFT(x)) == x on one very long signal
    (every CTA's strip lies inside the same signal), on thousands of very short ones (every tile
    ends at a signal end) and on a ragged in-between, with a loud and a quiet stretch per signal."""
    x = randn(shape, 71)
    x[..., : shape[1] // 3] *= 1e-3
    dev = x.to(DEV)
    stft = brv.STFT(**kw)
    y = stft.backward(stft(dev))[..., : shape[1]]
    n = y.shape[-1]                    # hop * (T - 1): a few samples short of the input when hop does not divide n_fft
    assert shape[1] - n < stft.hop_length
    err = (y - dev[..., :n]).abs()
    assert float(err.max()) < 2e-5, (kw, shape, float(err.max()))
    if shape[1] // 3 > 1200:           # frames that only see the quiet stretch: per-frame scales keep them exact
        quiet = err[..., : shape[1] // 3 - 600]
        assert float(quiet.max()) < 2e-8, (kw, shape, float(quiet.max()))
