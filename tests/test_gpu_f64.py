"""GPU parity tests for the float64 compute path (brv_stft_f64.cu): a float64 tensor gets double
precision arithmetic, as it does from torch.stft / torch.istft in the reference
(brever/modules/stft.py:59-138), checked against the float64 numpy oracle at 1e-11 and, for the
gradients, against torch autograd through the reference's own library calls in float64 and
torch.autograd.gradcheck."""
import numpy as np
import pytest
import torch

import brever_b200 as brv
from oracle import tf_oracle as O
from oracle import torch_port as P

from _util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-11          # double arithmetic over 64 .. 512-term sums

CASES = [
    dict(frame_length=512, hop_length=128),
    dict(frame_length=512, hop_length=256, normalized=False, scale_factor=0.3),
    dict(frame_length=400, hop_length=100, n_fft=512),
    dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=64, hop_length=16, window='hamming'),
]


def randn64(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)


@pytest.mark.parametrize('kw', CASES)
@pytest.mark.parametrize('shape', [(1, 100), (3, 4097), (2, 2, 20000)])
def test_f64_forward_matches_oracle(kw, shape):
    x = randn64(shape, 3)
    spec = brv.STFT(**kw)(x.to(DEV))
    assert spec.dtype == torch.complex128
    ref = O.stft(x.numpy(), **kw)
    assert spec.shape == ref.shape
    e = rel_err(spec.cpu().numpy(), ref)
    assert e[0] < TOL and e[1] < TOL, (kw, shape, e)


@pytest.mark.parametrize('kw', CASES)
@pytest.mark.parametrize('frames', [2, 9, 130])
@pytest.mark.parametrize('layout', ['bin_major', 'frame_major'])
def test_f64_inverse_matches_oracle(kw, frames, layout):
    stft = brv.STFT(**kw)
    g = torch.Generator().manual_seed(5)
    spec = torch.complex(torch.randn(3, stft.n_bins, frames, generator=g, dtype=torch.float64),
                         torch.randn(3, stft.n_bins, frames, generator=g, dtype=torch.float64))
    try:
        ref = O.istft(spec.numpy(), **kw)
    except RuntimeError:
        pytest.skip('NOLA')
    dev = spec.to(DEV)
    if layout == 'frame_major':
        dev = dev.transpose(1, 2).contiguous().transpose(1, 2)
    keep = dev.clone()
    y = stft.backward(dev)
    assert y.dtype == torch.float64 and torch.equal(dev, keep)
    e = rel_err(y.cpu().numpy(), ref)
    assert e[0] < TOL and e[1] < TOL, (kw, frames, layout, e)


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128, pad_mode='reflect'),
                                dict(frame_length=400, hop_length=100, n_fft=512, pad_mode='reflect'),
                                dict(frame_length=512, hop_length=128, center=False)])
def test_f64_padding_modes(kw):
    x = randn64((3, 4097), 13)
    xg = x.clone().to(DEV).requires_grad_(True)
    spec = brv.STFT(**kw)(xg)
    ref = O.stft(x.numpy(), **kw)
    e = rel_err(spec.detach().cpu().numpy(), ref)
    assert e[0] < TOL and e[1] < TOL, (kw, e)
    w = torch.complex(randn64(tuple(spec.shape), 14), randn64(tuple(spec.shape), 15))
    (spec * w.conj().to(DEV)).real.sum().backward()
    e = rel_err(xg.grad.cpu().numpy(), O.stft_grad(w.numpy(), x.shape[-1], **kw))
    assert e[0] < TOL and e[1] < TOL, ('gradient', kw, e)


def test_f64_round_trip_and_return_types():
    stft = brv.STFT(512, 128)
    x = randn64((4, 16000), 9).to(DEV)
    spec = stft(x)
    y = stft.backward(spec)
    assert float((y[..., :16000] - x).abs().max()) < 1e-12
    re, im = stft(x, return_type='real_imag')
    assert re.dtype == torch.float64 and torch.equal(re, spec.real) and torch.equal(im, spec.imag)
    mag, ph = stft(x, return_type='mag_phase')
    y2 = stft.backward((mag, ph), input_type='mag_phase')
    assert float((y2 - y).abs().max()) < 1e-12


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128),
                                dict(frame_length=256, hop_length=64, normalized=False, scale_factor=0.5),
                                dict(frame_length=510, hop_length=128, normalized=False)])
def test_f64_gradients_match_reference_autograd(kw):
    stft = brv.STFT(**kw)
    win = torch.from_numpy(O.get_window('hann', kw['frame_length']))
    x = randn64((3, 5000), 21)
    frames = stft.n_frames(5000)
    g = torch.Generator().manual_seed(22)
    wt = torch.complex(torch.randn(3, stft.n_bins, frames, generator=g, dtype=torch.float64),
                       torch.randn(3, stft.n_bins, frames, generator=g, dtype=torch.float64))
    xg = x.clone().to(DEV).requires_grad_(True)
    (stft(xg) * wt.to(DEV).conj()).real.sum().backward()
    xr = x.clone().requires_grad_(True)
    (P.stft(xr, win, **kw) * wt.conj()).real.sum().backward()
    e = rel_err(xg.grad.cpu().numpy(), xr.grad.numpy())
    assert e[0] < TOL and e[1] < TOL, ('d stft / dx', kw, e)

    v = randn64((3, stft.hop_length * (frames - 1)), 23)
    sg = wt.clone().to(DEV).requires_grad_(True)
    (stft.backward(sg) * v.to(DEV)).sum().backward()
    sr = wt.clone().requires_grad_(True)
    (P.istft(sr, win, **kw) * v).sum().backward()
    e = rel_err(sg.grad.cpu().numpy(), sr.grad.numpy())
    assert e[0] < TOL and e[1] < TOL, ('d istft / dX', kw, e)


def test_f64_gradcheck():
    stft = brv.STFT(frame_length=64, hop_length=16)
    x = randn64((2, 300), 31).to(DEV).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: torch.view_as_real(stft(t)), (x,), nondet_tol=0.0)
    g = torch.Generator().manual_seed(32)
    spec = torch.complex(torch.randn(2, 33, 12, generator=g, dtype=torch.float64),
                         torch.randn(2, 33, 12, generator=g, dtype=torch.float64)).to(DEV)
    # Im X[0] and Im X[N/2] do not reach the output (c2r): gradcheck sees exactly that
    spec.requires_grad_(True)
    assert torch.autograd.gradcheck(lambda s: stft.backward(s), (spec,), nondet_tol=0.0)


def test_f64_float32_inputs_still_take_the_tensor_core_path():
    lib = brv._lib.lib()
    stft = brv.STFT(512, 128)
    x = torch.randn(4, 16000, device=DEV)
    stft(x)
    n0 = lib.brv_launch_count()
    spec = stft(x)
    assert lib.brv_launch_count() - n0 == 1 and spec.dtype == torch.complex64
