"""The tcgen05 STFT path against the generic CUDA-core path, the float64 oracle
and the golden vectors, with the measured error printed (pytest -s / log)."""
import numpy as np
import pytest
import torch

import brever_b200 as brv
from brever_b200 import _lib
from oracle import tf_oracle as O

from _util import assert_parity, crandn, golden, randn, rel_err, synthetic_mixture

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def cpu(t):
    return t.detach().cpu().numpy()


class generic_path:
    def __enter__(self):
        self.prev = _lib.lib().brv_set_force_generic(1)

    def __exit__(self, *a):
        _lib.lib().brv_set_force_generic(self.prev)


class tc_variant:
    """0: default kernels, 1: dense contraction only, 2 / 3: folded forward forced to one tile
    per CTA / to the persistent two-pass kernel."""

    def __init__(self, variant):
        self.variant = variant

    def __enter__(self):
        self.prev = _lib.lib().brv_set_tc_variant(self.variant)

    def __exit__(self, *a):
        _lib.lib().brv_set_tc_variant(self.prev)


CASES = [
    dict(frame_length=512, hop_length=128),
    dict(frame_length=512, hop_length=256),
    dict(frame_length=256, hop_length=128, normalized=False),
    dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5,
         scale_factor=0.15),
    dict(frame_length=400, hop_length=100, n_fft=512),
    dict(frame_length=512, hop_length=100, window='hamming'),
    dict(frame_length=1024, hop_length=256),
    dict(frame_length=64, hop_length=16),
]


@pytest.mark.parametrize('kw', CASES)
@pytest.mark.parametrize('samples', [100, 4097, 40000])
def test_tensorcore_forward_matches_generic_and_oracle(kw, samples, capsys):
    x = randn((3, samples), 77)
    x[1] *= 1e-3          # quiet signal: per-frame scaling keeps its relative accuracy
    x[2] *= 300.0         # loud signal
    stft = brv.STFT(**kw)
    tc = stft(x.to(DEV))
    with generic_path():
        gen = stft(x.to(DEV))
    ref = O.stft(x.numpy(), **{k: v for k, v in kw.items()})
    assert tc.shape == gen.shape == ref.shape
    assert tc.stride() == gen.stride()
    worst = 0.0
    for i in range(3):
        e_tc = rel_err(cpu(tc[i]), ref[i])
        e_gen = rel_err(cpu(gen[i]), ref[i])
        worst = max(worst, e_tc[0])
        assert e_tc[0] < 1e-4 and e_tc[1] < 1e-4, (kw, i, e_tc)
        assert e_gen[0] < 1e-4, (kw, i, e_gen)
    with capsys.disabled():
        print(f'\n[tc-accuracy] {kw} S={samples}: max-rel {worst:.2e}')


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128),
                                dict(frame_length=512, hop_length=256),
                                dict(frame_length=256, hop_length=128, normalized=False),
                                dict(frame_length=128, hop_length=32),
                                dict(frame_length=384, hop_length=96, compression_factor=0.5,
                                     scale_factor=0.15),
                                # n_fft = 4Q - 2: no Nyquist bin, no n = Q column, unaligned centre pad
                                dict(frame_length=510, hop_length=128, normalized=False,
                                     compression_factor=0.5, scale_factor=0.15),
                                dict(frame_length=510, hop_length=128),
                                dict(frame_length=254, hop_length=64, window='hamming'),
                                dict(frame_length=200, hop_length=50, n_fft=254)])
@pytest.mark.parametrize('shape', [(1, 100), (3, 4097), (400, 20000), (37, 64000)])
def test_pipelined_forward_matches_single_tile_kernel_and_oracle(kw, shape):
    """The persistent two-pass forward (several tiles per CTA, TMEM halves handed back
    and forth) against the one-tile-per-CTA kernel and the float64 oracle."""
    x = randn(shape, 5)
    x[::3] *= 1e-3
    x[1::3] *= 200.0
    stft = brv.STFT(**kw)
    with tc_variant(3):
        new = stft(x.to(DEV))
        again = stft(x.to(DEV))
    with tc_variant(2):
        old = stft(x.to(DEV))
    assert new.shape == old.shape and new.stride() == old.stride()
    assert torch.equal(new, again)                     # deterministic
    for i in range(0, shape[0], max(1, shape[0] // 7)):
        e = rel_err(cpu(new[i]), cpu(old[i]))
        assert e[0] < 2e-6, (kw, shape, i, e)
        ref = O.stft(x[i].numpy(), **kw)
        e = rel_err(cpu(new[i]), ref)
        assert e[0] < 1e-4 and e[1] < 1e-4, (kw, shape, i, e)


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128),
                                dict(frame_length=512, hop_length=256, scale_factor=0.3),
                                dict(frame_length=256, hop_length=128, normalized=False),
                                dict(frame_length=400, hop_length=128, n_fft=512),
                                dict(frame_length=510, hop_length=128, normalized=False),   # n_fft = 4Q - 2
                                dict(frame_length=254, hop_length=64)])
@pytest.mark.parametrize('shape', [(2, 3), (3, 130), (160, 157), (64, 501)])
@pytest.mark.parametrize('variant', [2, 3])
def test_tensorcore_istft_gradient(kw, shape, variant):
    """d STFT.backward / dX on the folded tcgen05 forward kernel (gy / envelope staged as the
    signal, Hermitian bin weights) against the generic float64-accumulating path and, on a
    few signals, torch autograd through the reference's own library calls."""
    from oracle import torch_port as P
    n_sig, frames = shape
    stft = brv.STFT(**kw)
    spec = crandn((n_sig, stft.n_bins, frames), 31)
    v = randn((n_sig, stft.hop_length * (frames - 1)), 32)
    v[::2] *= 1e-3
    lib = _lib.lib()

    def grad():
        sg = spec.clone().to(DEV).requires_grad_(True)
        n0 = lib.brv_launch_count()
        (stft.backward(sg) * v.to(DEV)).sum().backward()
        return sg.grad, lib.brv_launch_count() - n0

    with tc_variant(variant):
        g_tc, launches = grad()
    assert launches == 2                      # one inverse kernel + one gradient kernel
    with generic_path():
        g_gen, _ = grad()
    assert g_tc.shape == g_gen.shape == spec.shape
    for i in range(0, n_sig, max(1, n_sig // 5)):
        e = rel_err(cpu(g_tc[i]), cpu(g_gen[i]))
        assert e[0] < 2e-5 and e[1] < 2e-5, (kw, shape, i, e)
    assert float(g_tc[:, 0].imag.abs().max()) == 0       # DC / Nyquist carry no imaginary gradient
    assert float(g_tc[:, -1].imag.abs().max()) == 0
    win = torch.from_numpy(O.get_window('hann', kw['frame_length']))
    sr = spec[:2].clone().requires_grad_(True)
    (P.istft(sr, win, **kw) * v[:2]).sum().backward()
    assert_parity(cpu(g_tc[:2]), sr.grad.numpy(), 1e-4, 'd istft / dX')


@pytest.mark.parametrize('kw', [dict(frame_length=512, hop_length=128),
                                dict(frame_length=512, hop_length=256, scale_factor=0.3),
                                dict(frame_length=256, hop_length=128, normalized=False),
                                dict(frame_length=400, hop_length=128, n_fft=512),
                                dict(frame_length=510, hop_length=128, normalized=False),   # n_fft = 4Q - 2
                                dict(frame_length=254, hop_length=64)])
@pytest.mark.parametrize('shape', [(2, 100), (3, 4097), (160, 20000), (64, 64000)])
def test_tensorcore_stft_gradient(kw, shape):
    """d STFT.forward / dx on the folded tcgen05 inverse kernel (all bins weigh 1, no
    envelope, cut to the input length) against the generic path and reference autograd."""
    from oracle import torch_port as P
    n_sig, samples = shape
    stft = brv.STFT(**kw)
    frames = stft.n_frames(samples)
    x = randn(shape, 41)
    wt = crandn((n_sig, stft.n_bins, frames), 42)
    wt[::2] *= 1e-2
    lib = _lib.lib()

    def grad():
        xg = x.clone().to(DEV).requires_grad_(True)
        n0 = lib.brv_launch_count()
        (stft(xg) * wt.to(DEV).conj()).real.sum().backward()
        return xg.grad, lib.brv_launch_count() - n0

    g_tc, launches = grad()
    assert launches == 2                      # one forward kernel + one gradient kernel
    with generic_path():
        g_gen, _ = grad()
    assert g_tc.shape == g_gen.shape == x.shape
    for i in range(0, n_sig, max(1, n_sig // 5)):
        e = rel_err(cpu(g_tc[i]), cpu(g_gen[i]))
        assert e[0] < 2e-5 and e[1] < 2e-5, (kw, shape, i, e)
    win = torch.from_numpy(O.get_window('hann', kw['frame_length']))
    xr = x[:2].clone().requires_grad_(True)
    (P.stft(xr, win, **kw) * wt[:2].conj()).real.sum().backward()
    assert_parity(cpu(g_tc[:2]), xr.grad.numpy(), 1e-4, 'd stft / dx')


def test_tensorcore_is_actually_used():
    """The default path must launch the tcgen05 kernel (count launches)."""
    lib = _lib.lib()
    stft = brv.STFT(512, 128)
    x = torch.randn(4, 16000, device=DEV)
    stft(x)
    n0 = lib.brv_launch_count()
    stft(x)
    assert lib.brv_launch_count() - n0 == 1


def test_tensorcore_golden_and_roundtrip():
    g = golden()
    x = randn((4096,), 42)
    for hop in (256, 128):
        stft = brv.STFT(512, hop)
        spec = stft(x.to(DEV))
        assert_parity(cpu(spec), g[f'rt_h{hop}_c1.0_s1.0_n1_o1_spec'], 1e-4)
        y = stft.backward(spec).cpu()
        err = float((y - x).abs().max())
        print(f'[tc-accuracy] round trip hop {hop}: max abs err {err:.2e}')
        assert err < 1e-4 * float(x.abs().max())
        assert torch.allclose(x, y, rtol=2e-3, atol=2e-6)


def test_tensorcore_nan_and_inf_stay_local():
    stft = brv.STFT(512, 128)
    x = torch.randn(3, 20000)
    x[0, 7000] = float('nan')
    x[1, 9000] = float('inf')
    spec = stft(x.to(DEV))
    t = np.arange(spec.shape[-1])
    for row, i in ((0, 7000), (1, 9000)):
        bad = (~torch.isfinite(spec[row].real)).any(0).cpu().numpy()
        expect = (t * 128 - 256 <= i) & (i < t * 128 - 256 + 512)
        assert np.array_equal(bad, expect), row
    assert torch.isfinite(spec[2].real).all()


def test_tensorcore_baseline_size_cfg2():
    mix, _ = synthetic_mixture((64, 64000), 1000)
    stft = brv.STFT(512, 128)
    tc = stft(mix.to(DEV))
    with generic_path():
        gen = stft(mix.to(DEV))
    e = rel_err(cpu(tc), cpu(gen))
    print(f'[tc-accuracy] cfg2 tc vs generic: {e}')
    assert e[0] < 2e-5


@pytest.mark.parametrize('kw', CASES)
@pytest.mark.parametrize('frames', [1, 9, 200])
@pytest.mark.parametrize('layout', ['bin_major', 'frame_major'])
def test_tensorcore_inverse_matches_generic_and_oracle(kw, frames, layout, capsys):
    from _util import crandn
    stft = brv.STFT(**kw)
    spec = crandn((3, stft.n_bins, frames), 91)
    spec[1] *= 1e-3
    spec[2] *= 300.0
    dev = spec.to(DEV)
    if layout == 'frame_major':
        dev = dev.transpose(1, 2).contiguous().transpose(1, 2)
    try:
        ref = O.istft(spec.numpy(), **kw)
    except RuntimeError:
        with pytest.raises(RuntimeError):
            stft.backward(dev)
        return
    tc = stft.backward(dev)
    with generic_path():
        gen = stft.backward(dev)
    assert tc.shape == gen.shape == ref.shape
    worst = 0.0
    for i in range(3):
        e_tc = rel_err(cpu(tc[i]), ref[i])
        worst = max(worst, e_tc[0])
        assert e_tc[0] < 1e-4 and e_tc[1] < 1e-4, (kw, i, e_tc)
        assert rel_err(cpu(gen[i]), ref[i])[0] < 1e-4
    with capsys.disabled():
        print(f'\n[tc-accuracy] inverse {kw} T={frames} {layout}: max-rel {worst:.2e}')


def test_tensorcore_roundtrip_baseline_sizes():
    for shape, kw in [((64, 64000), dict(frame_length=512, hop_length=128)),
                      ((8, 128000), dict(frame_length=510, hop_length=128, normalized=False,
                                         compression_factor=0.5, scale_factor=0.15))]:
        mix, _ = synthetic_mixture(shape, 1000)
        stft = brv.STFT(**kw)
        y = stft.backward(stft(mix.to(DEV)))[..., :shape[-1]]
        e = rel_err(cpu(y), mix.numpy())
        print(f'[tc-accuracy] round trip {kw}: {e}')
        assert e[0] < 1e-4 and e[1] < 1e-4


# ---- folded inverse with the fused overlap-add (brv_stft_fold.cu) ----------------
FOLD_INV_CASES = [
    dict(frame_length=512, hop_length=128),                       # HQ = 1 (DCCRN)
    dict(frame_length=512, hop_length=256),                       # HQ = 2 (FFNN)
    dict(frame_length=256, hop_length=128, normalized=False),     # HQ = 2 (TF-GridNet)
    dict(frame_length=256, hop_length=64),                        # HQ = 1
    dict(frame_length=256, hop_length=256, window=None),          # HQ = 4 (no overlap)
    dict(frame_length=128, hop_length=32, window='hamming'),      # Q = 32
    dict(frame_length=384, hop_length=96),                        # Q = 96
    dict(frame_length=400, hop_length=128, n_fft=512),            # window shorter than n_fft
    dict(frame_length=512, hop_length=128, compression_factor=0.5, scale_factor=0.15,
         normalized=False),                                        # SGMSE-style decompression
    dict(frame_length=510, hop_length=128, compression_factor=0.5, scale_factor=0.15,
         normalized=False),                                        # SGMSE (cfg4): n_fft = 4Q - 2
    dict(frame_length=510, hop_length=128),
    dict(frame_length=254, hop_length=64, window='hamming'),
    dict(frame_length=126, hop_length=32, normalized=False),
]


@pytest.mark.parametrize('kw', FOLD_INV_CASES)
@pytest.mark.parametrize('frames', [2, 4, 127, 129, 131, 260, 501])
@pytest.mark.parametrize('layout', ['bin_major', 'frame_major', 'strided'])
def test_folded_inverse_matches_oracle(kw, frames, layout):
    """Tile seams (128-frame tiles overlapping by R - 1 frames), warp seams (spill
    slots) and all three load mappings against the float64 oracle."""
    from _util import crandn
    stft = brv.STFT(**kw)
    spec = crandn((3, stft.n_bins, frames), 1234 + frames)
    spec[1] *= 1e-3
    spec[2, :, frames // 2:] *= 1e3       # loud second half: per-frame scales differ
    if layout == 'bin_major':
        dev = spec.to(DEV)
    elif layout == 'frame_major':
        dev = spec.to(DEV).transpose(1, 2).contiguous().transpose(1, 2)
    else:   # every other frame of a wider buffer, bins not contiguous either
        wide = torch.zeros((3, stft.n_bins + 3, 2 * frames), dtype=torch.complex64, device=DEV)
        wide[:, 1:1 + stft.n_bins, ::2] = spec.to(DEV)
        dev = wide[:, 1:1 + stft.n_bins, ::2]
    try:
        ref = O.istft(spec.numpy(), **kw)
    except RuntimeError:
        with pytest.raises(RuntimeError):
            stft.backward(dev)
        return
    lib = _lib.lib()
    n0 = lib.brv_launch_count()
    got = stft.backward(dev)
    assert lib.brv_launch_count() - n0 == 1, 'expected the single fused kernel'
    assert got.shape == ref.shape
    for i in range(3):
        e = rel_err(cpu(got[i]), ref[i])
        assert e[0] < 1e-4 and e[1] < 1e-4, (kw, frames, layout, i, e)
    prev = lib.brv_set_tc_variant(1)       # dense contraction + separate overlap-add
    try:
        dense = stft.backward(dev)
    finally:
        lib.brv_set_tc_variant(prev)
    assert rel_err(cpu(got), cpu(dense))[0] < 2e-5


def test_folded_inverse_is_deterministic_and_nan_local():
    stft = brv.STFT(512, 128)
    from _util import crandn
    spec = crandn((2, 257, 300), 5).to(DEV)
    a = stft.backward(spec)
    b = stft.backward(spec)
    assert torch.equal(a, b)
    spec[0, 40, 150] = complex(float('nan'), 0.0)
    y = stft.backward(spec)
    bad = ~torch.isfinite(y[0])
    idx = bad.nonzero().flatten().cpu().numpy()
    # frame 150 covers trimmed samples [150*128 - 256, 150*128 + 256)
    assert idx.min() == 150 * 128 - 256 and idx.max() == 150 * 128 + 255
    assert torch.isfinite(y[1]).all()
