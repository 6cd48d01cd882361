"""GPU parity tests of the round-2 additions against reference-generated goldens: STFT padding /
centre modes, MANNER's multi-resolution STFT loss, the batched on-device transforms, the
scale-invariant MultiResYuLoss gradient."""
import numpy as np
import pytest
import torch

import brever_b200 as brv

from oracle import tf_oracle as O

from _util import assert_parity, crandn, golden, randn

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def cpu(t):
    return t.detach().cpu().numpy()


PAD_CASES = [('reflect_512_128', dict(frame_length=512, hop_length=128, pad_mode='reflect'), (3, 3001), 700),
             ('reflect_256_64', dict(frame_length=256, hop_length=64, pad_mode='reflect', normalized=False), (2, 2, 1500), 701),
             ('nocenter_512_128', dict(frame_length=512, hop_length=128, center=False), (3, 3001), 702),
             ('reflect_nocenter_400', dict(frame_length=400, hop_length=100, n_fft=512, center=False, pad_mode='reflect'), (2, 2777), 703)]


@pytest.mark.parametrize('tag,kw,shape,seed', PAD_CASES)
def test_stft_padding_modes(tag, kw, shape, seed):
    """stft.py:59-77,140-144: pad_mode='reflect' (tail + centre mirrored) and center=False."""
    g = golden()
    stft = brv.STFT(**kw)
    x = randn(shape, seed)
    spec = stft(x.to(DEV))
    assert tuple(spec.shape) == g[f'stft_{tag}'].shape
    assert_parity(cpu(spec), g[f'stft_{tag}'], 1e-4, tag)
    w = crandn(tuple(spec.shape), seed + 50).to(DEV)
    xg = x.clone().to(DEV).requires_grad_(True)
    sg = stft(xg)
    (sg.real * w.real + sg.imag * w.imag).sum().backward()
    assert_parity(cpu(xg.grad), g[f'stft_{tag}_grad'], 1e-4, tag + ' gradient')
    if kw.get('center', True):
        assert_parity(cpu(stft.backward(spec)), g[f'stft_{tag}_back'], 1e-4, tag + ' inverse')
    else:
        with pytest.raises(NotImplementedError):
            stft.backward(spec)
    with pytest.raises(NotImplementedError):
        brv.STFT(pad_mode='replicate')


@pytest.mark.parametrize('tag,kw', [('def', {}), ('small', dict(fft_sizes=[512, 256], hop_sizes=[128, 64],
                                                              win_lengths=[512, 200], factor_sc=0.5, factor_mag=1.0))])
def test_manner_multi_resolution_stft_loss(tag, kw):
    """models/manner/stft_loss.py:22-151: values and gradient (1024 / 2048-point transforms on
    the dense tcgen05 path, 512 / 256 on the folded one; reflect padding; clamp-sqrt)."""
    g = golden()
    mx, my = 0.1 * randn((3, 8000), 710), 0.1 * randn((3, 8000), 711)
    my = 0.7 * mx + 0.3 * my
    my[2, 5000:] = 0.0
    crit = brv.manner.MultiResolutionSTFTLoss(**kw)
    e = mx.clone().to(DEV).requires_grad_(True)
    sc, mag = crit(e, my.to(DEV))
    assert sc.shape == (3,) and mag.shape == (3,)
    assert np.allclose(cpu(sc), g[f'manner_{tag}_sc'], rtol=1e-4)
    assert np.allclose(cpu(mag), g[f'manner_{tag}_mag'], rtol=1e-4)
    wsc, wmag = torch.tensor([1.0, -0.5, 2.0], device=DEV), torch.tensor([0.3, 1.5, -1.0], device=DEV)
    ((sc * wsc).sum() + (mag * wmag).sum()).backward()
    # the log-magnitude term's gradient is sign(|X| - |Y|) / |X|: wherever float32 rounding decides
    # the sign (near-ties, the clamped silence of item 2) reference and kernel legitimately differ
    assert_parity(cpu(e.grad), g[f'manner_{tag}_grad'], 1e-3, 'manner gradient')


def _tbatch():
    batch = 0.05 * randn((3, 2, 2, 4000), 720)
    blen = [4000, 3000, 2345]
    for i, n in enumerate(blen):
        batch[i, ..., n:] = 0
    return batch, blen


@pytest.mark.parametrize('tag,kw', [('def', {}), ('s3d2', dict(stacks=3, decimation=2))])
def test_ffnn_transform_batched(tag, kw):
    """training.py:336-338 + data.py:408-491: the per-utterance transform loop + collate."""
    g = golden()
    batch, blen = _tbatch()
    front = brv.ffnn.FFNNFrontEnd(**kw)
    out, frames = brv.transform_batched('ffnn', batch.to(DEV), torch.tensor(blen), front=front)
    assert frames.tolist() == g[f'ffnn_tbatch_{tag}_len'].tolist()
    assert tuple(out.shape) == g[f'ffnn_tbatch_{tag}'].shape
    n_feat = front.input_size
    assert_parity(cpu(out[:, :n_feat]), g[f'ffnn_tbatch_{tag}'][:, :n_feat], 1e-4, 'features')
    assert_parity(cpu(out[:, n_feat:]), g[f'ffnn_tbatch_{tag}'][:, n_feat:], 1e-4, 'labels')
    for i, t in enumerate(frames.tolist()):
        assert float(out[i, :, t:].abs().max()) == 0 if t < out.shape[-1] else True
        one = front.transform(batch[i, ..., :blen[i]].to(DEV))          # per-utterance path
        assert_parity(cpu(out[i, :, :t]), cpu(one), 2e-6, 'batched == per item')


def test_sgmse_transform_batched():
    g = golden()
    batch, blen = _tbatch()
    stft = brv.STFT(frame_length=512, hop_length=128, window='hann', compression_factor=0.5,
                    scale_factor=0.15, normalized=False)
    out, frames = brv.transform_batched('sgmsep', batch.to(DEV), torch.tensor(blen), stft=stft)
    assert frames.tolist() == g['sgmse_tbatch_len'].tolist()
    assert tuple(out.shape) == g['sgmse_tbatch'].shape
    assert_parity(cpu(out), g['sgmse_tbatch'], 1e-4, 'sgmse transform')
    with pytest.raises(ValueError):
        brv.transform_batched('dccrn', batch.to(DEV), torch.tensor(blen))


def test_multiresyu_scale_invariant_gradient():
    """criterion.py:207-226 with scale_invariant=True: gradient through the scaling factor."""
    g = golden()
    B, S, L = 5, 3, 2000
    est, ref = randn((B, S, L), 600), randn((B, S, L), 601)
    est = ref.roll(1, 1) * 0.7 + 0.3 * est
    lengths = torch.from_numpy(g['crit_lengths'])
    weight = torch.from_numpy(g['crit_mse_weight'])
    crit = brv.MultiResYuLoss(scale_invariant=True)
    e = est.clone().to(DEV).requires_grad_(True)
    out = crit(e, ref.to(DEV), lengths)
    assert np.allclose(cpu(out), g['crit_mry_si'], rtol=2e-5)
    (out * weight.to(DEV)).sum().backward()
    assert_parity(cpu(e.grad), g['crit_mry_si_grad'], 1e-4, 'SI gradient')


@pytest.mark.parametrize('normalized', [True, False])
def test_conv_stft_gradients(normalized):
    """ConvSTFT is differentiable in the reference (F.conv1d / F.conv_transpose1d, stft.py:244-300):
    both directions against torch autograd through that formulation in float64 on the CPU."""
    import torch.nn.functional as F
    L, H, s = 256, 64, 0.7
    conv = brv.ConvSTFT(frame_length=L, hop_length=H, scale_factor=s, normalized=normalized)
    filters = conv.filters.double()
    dim = L // 2 + 1
    nf2 = 1.0 if normalized else conv._normalization_factor ** 2
    x = randn((3, 2000), 41)
    w = crandn((3, dim, conv.n_frames(2000)), 42)

    def ref_forward(xd):
        out = F.conv1d(conv.pad(xd).unsqueeze(1), filters, stride=H) * s
        return torch.complex(out[:, :dim], out[:, dim:])

    def ref_backward(X):
        z = torch.cat([X.real, X.imag], dim=-2) / s
        y = F.conv_transpose1d(z, filters, stride=H) / nf2
        return y[:, 0, L - H:-(L - H)]

    xr = x.double().requires_grad_(True)
    (ref_forward(xr) * w.conj().to(torch.complex128)).real.sum().backward()
    xg = x.clone().to(DEV).requires_grad_(True)
    (conv(xg) * w.conj().to(DEV)).real.sum().backward()
    assert_parity(cpu(xg.grad), xr.grad.numpy(), 1e-4, 'd ConvSTFT / dx')

    X = crandn((3, dim, 40), 43)
    v = randn((3, 41 * H - L), 44)
    Xr = X.to(torch.complex128).requires_grad_(True)
    (ref_backward(Xr) * v.double()).sum().backward()
    Xg = X.clone().to(DEV).requires_grad_(True)
    (conv.backward(Xg) * v.to(DEV)).sum().backward()
    assert_parity(cpu(Xg.grad), Xr.grad.numpy(), 1e-4, 'd ConvSTFT.backward / dX')
    with pytest.raises(NotImplementedError):
        c2 = brv.ConvSTFT(frame_length=L, hop_length=H, compression_factor=0.5)
        c2(x.clone().to(DEV).requires_grad_(True))


def test_istft_complex32_input():
    """AMP: TF-GridNet hands torch.complex(half, half) to STFT.backward; computed
    in float32, returned as float16."""
    stft = brv.STFT(256, 128, normalized=False)
    spec = crandn((2, 129, 60), 51).to(DEV)
    half = torch.complex(spec.real.half(), spec.imag.half())
    assert half.dtype == torch.complex32
    y = stft.backward(half)
    assert y.dtype == torch.float16
    ref = stft.backward(torch.complex(half.real.float(), half.imag.float()))
    assert torch.allclose(y.float(), ref, atol=2e-3 * float(ref.abs().max()))


BENCH_SHAPES = [
    ('cfg3', None, (256, 1, 64000)),
    ('cfg4', dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15),
     (128, 128000)),
    ('cfg5', dict(frame_length=256, hop_length=128, normalized=False), (1024, 2, 64000)),
]


@pytest.mark.parametrize('name,kw,shape', BENCH_SHAPES)
def test_exact_bench_shapes(name, kw, shape):
    """The tensors bench.py times (same generator, same seeds): >= 8 signals spread over the batch
    against the float64 oracle, a checksum of every row against the generic path, and the round
    trip -- the dispatch (strip kernel, strip scheduling, persistent tile order) depends on the
    launch size, so reduced batches do not cover it."""
    from _util import synthetic_mixture
    from brever_b200 import _lib
    from oracle import tf_oracle as O
    mix, fg = synthetic_mixture(shape, 1000 + int(name[-1]))
    if kw is None:                                      # cfg3: criterion only
        lengths = torch.full((shape[0],), shape[-1])
        out = brv.sisnr(mix.to(DEV), fg.to(DEV), lengths)
        rows = np.linspace(0, shape[0] - 1, 8).astype(int)
        ref, _ = O.sisnr(mix[rows].numpy(), fg[rows].numpy(), lengths[rows].numpy())
        assert np.allclose(cpu(out[rows]), ref, atol=1e-4)
        return
    stft = brv.STFT(**kw)
    x = mix.to(DEV)
    spec = stft(x)
    flat = x.reshape(-1, x.shape[-1])
    fspec = spec.reshape(-1, *spec.shape[-2:])
    rows = np.linspace(0, flat.shape[0] - 1, 8).astype(int)
    for r in rows:
        assert_parity(cpu(fspec[r]), O.stft(cpu(flat[r]), **kw), 1e-4, f'{name} stft row {r}')
    # checksum of every row against the generic (float64-accumulating SIMT) path, in chunks
    prev = _lib.lib().brv_set_force_generic(1)
    try:
        for start in range(0, flat.shape[0], 256):
            gen = stft(flat[start:start + 256])
            a = fspec[start:start + 256]
            num = (a - gen).abs().amax(dim=(-2, -1))
            den = gen.abs().amax(dim=(-2, -1))
            assert float((num / den).max()) < 1e-4, (name, start)
            del gen
    finally:
        _lib.lib().brv_set_force_generic(prev)
    sel = spec[:, 0] if spec.ndim == 4 else spec
    y = stft.backward(sel)[..., :shape[-1]]
    xs = x[:, 0] if x.ndim == 3 else x
    err = (y - xs).abs().amax(-1) / xs.abs().amax(-1)
    assert float(err.max()) < 2e-4, (name, float(err.max()))
    for r in np.linspace(0, sel.shape[0] - 1, 8).astype(int):
        assert_parity(cpu(y[r]), O.istft(cpu(sel[r]), **kw)[..., :shape[-1]], 1e-4, f'{name} istft row {r}')


def test_channel_mean_mask_and_accumulate():
    """FFNN._enhance's `x.mean(1) * mask` (ffnn.py:107-110) and the running-loss update in one launch each."""
    spec = crandn((3, 2, 257, 40), 61).to(DEV)
    mask = randn((3, 257, 40), 62).abs().to(DEV)
    out = brv.ffnn.channel_mean(spec, mask)
    ref = spec.mean(1) * mask
    assert out.shape == ref.shape and out.transpose(1, 2).is_contiguous()
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-7)
    assert torch.allclose(brv.ffnn.channel_mean(spec.transpose(2, 3).contiguous().transpose(2, 3)), spec.mean(1),
                          rtol=1e-6, atol=1e-7)
    # the layout pair of the real call (frame-major STFT output, bin-major mask): tiled kernel
    fm = crandn((3, 2, 257, 45), 64).to(DEV).transpose(2, 3).contiguous().transpose(2, 3)
    m2 = randn((3, 257, 45), 65).abs().to(DEV)
    out2 = brv.ffnn.channel_mean(fm, m2)
    assert out2.transpose(1, 2).is_contiguous()
    assert torch.allclose(out2, fm.mean(1) * m2, rtol=1e-6, atol=1e-7)
    total = torch.zeros((), device=DEV)
    v = randn((37,), 63).to(DEV)
    brv.ffnn.accumulate_mean(total, v)
    brv.ffnn.accumulate_mean(total, v)
    assert abs(float(total) - 2 * float(v.mean())) < 1e-6


@pytest.mark.parametrize('kw', [dict(frame_length=400, hop_length=100),
                                dict(frame_length=320, hop_length=160, normalized=False, scale_factor=0.5),
                                dict(frame_length=250, hop_length=125, window='hamming'),
                                dict(frame_length=64, hop_length=24, compression_factor=0.5, scale_factor=0.15),
                                dict(frame_length=101, hop_length=50)])
def test_conv_stft_arbitrary_sizes(kw):
    """ConvSTFT (stft.py:201-319) at sizes the folded tensor-core kernels do not cover: analysis,
    synthesis and, without compression, both gradients (the synthesis is the analysis' adjoint)."""
    conv = brv.ConvSTFT(**kw)
    x = randn((3, 5000), 61)
    spec = conv(x.to(DEV))
    ref = O.conv_stft(x.numpy(), **kw)
    assert spec.shape == ref.shape and spec.dtype == torch.complex64
    assert_parity(cpu(spec), ref, 2e-5, 'ConvSTFT')
    y = conv.backward(spec)
    assert_parity(cpu(y), O.conv_istft(ref, **kw), 2e-5, 'ConvSTFT.backward')
    assert_parity(cpu(conv.backward(spec.contiguous())), cpu(y), 1e-6, 'layouts')
    if kw.get('compression_factor', 1) != 1:
        return
    # <A x, w> == <x, A^H w>: the gradient of the analysis is the synthesis kernel and vice versa
    w = crandn(tuple(spec.shape), 62)
    xg = x.clone().to(DEV).requires_grad_(True)
    (conv(xg) * w.conj().to(DEV)).real.sum().backward()
    v = randn(tuple(y.shape), 63)
    Xg = w.clone().to(DEV).requires_grad_(True)
    (conv.backward(Xg) * v.to(DEV)).sum().backward()
    eps = 1e-3
    d = randn((3, 5000), 64)
    f = lambda t: float((conv(t.to(DEV)) * w.conj().to(DEV)).real.sum())   # noqa: E731
    num = (f(x + eps * d) - f(x - eps * d)) / (2 * eps)
    ana = float((xg.grad.cpu() * d).sum())
    assert abs(num - ana) <= 2e-3 * max(1.0, abs(ana)), (kw, num, ana)
    D = crandn(tuple(spec.shape), 65)
    h = lambda S: float((conv.backward(S.to(DEV)) * v.to(DEV)).sum())      # noqa: E731
    num = (h(w + eps * D) - h(w - eps * D)) / (2 * eps)
    ana = float((Xg.grad.cpu().real * D.real + Xg.grad.cpu().imag * D.imag).sum())
    assert abs(num - ana) <= 2e-3 * max(1.0, abs(ana)), (kw, num, ana)
