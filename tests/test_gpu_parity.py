"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden
vectors produced by the reference, against the CPU oracle on the same seeded
inputs, and — at BASELINE.json's full sizes — through size-independent
properties (round trips, linearity, batched == single).

Tolerance (north_star): spectra / features / waveforms within 1e-4 relative
(max-norm relative to max|ref| AND Frobenius), shapes / strides / frame counts
bit-exact.  Criteria are compared in dB with an absolute tolerance.
"""
import itertools

import numpy as np
import pytest
import torch

import brever_b200 as brv
from brever_b200 import _lib
from oracle import tf_oracle as O
from oracle import torch_port as P

from _util import (STFT_SHAPE_CASES, assert_parity, crandn, golden, randn,
                   rel_err, synthetic_mixture)

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-4

COMBOS = list(itertools.product([256, 128], [1.0, 0.5], [1.0, 0.15],
                                [False, True], [False, True]))


def cpu(t):
    return t.detach().cpu().numpy()


# --------------------------------------------------------------------------- #
# STFT / iSTFT                                                                #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize('hop,c,s,normalized,onesided', COMBOS)
def test_reference_stft_test_cases(hop, c, s, normalized, onesided):
    """The reference's own test (tests/test_modules.py:300-326), same input,
    same acceptance tolerances, plus value parity with its recorded outputs."""
    g = golden()
    kw = dict(frame_length=512, hop_length=hop, compression_factor=c,
              scale_factor=s, normalized=normalized, onesided=onesided)
    key = f'rt_h{hop}_c{c}_s{s}_n{int(normalized)}_o{int(onesided)}'
    stft = brv.STFT(**kw)
    x = randn((4096,), 42)
    spec = stft(x.to(DEV))
    assert spec.dtype == torch.complex64
    if key + '_spec' in g:
        assert_parity(cpu(spec), g[key + '_spec'], TOL, key)
    assert_parity(cpu(spec), O.stft(x.numpy(), **kw), TOL, key + ' oracle')
    y = stft.backward(spec).cpu()
    assert y.shape == x.shape
    # default (tensor-core) path: north_star tolerance 1e-4 relative; measured
    # ~2e-6 (fp32 accumulation over n_fft/16 tensor-core steps)
    assert_parity(y.numpy(), x.numpy(), TOL, key + ' round trip')
    assert torch.allclose(x, y, rtol=2e-3, atol=2e-5)
    assert_parity(y.numpy(), g[key + '_back'], TOL, key + ' back')
    # generic CUDA-core path (float64 accumulation): the reference's own
    # acceptance tolerances, verbatim (tests/test_modules.py:325-326)
    prev = _lib.lib().brv_set_force_generic(1)
    try:
        y = stft.backward(stft(x.to(DEV))).cpu()
    finally:
        _lib.lib().brv_set_force_generic(prev)
    assert torch.allclose(x, y, rtol=0, atol=1e-6)
    assert torch.allclose(x, y, rtol=2e-3, atol=0)


@pytest.mark.parametrize('i', range(len(STFT_SHAPE_CASES)))
def test_shapes_strides_and_values(i):
    g = golden()
    S, L, H, nfft, win = STFT_SHAPE_CASES[i]
    stft = brv.STFT(frame_length=L, hop_length=H, window=win, n_fft=nfft)
    x = randn((2, S), 100 + i)
    spec = stft(x.to(DEV))
    ref = g[f'shape{i}_spec']
    assert tuple(spec.shape) == ref.shape                      # bit-exact
    assert tuple(spec.stride()) == tuple(g[f'shape{i}_strides'])  # frame-major
    assert_parity(cpu(spec), ref, TOL, f'shape{i}')
    back = stft.backward(torch.from_numpy(ref).to(DEV))
    assert tuple(back.shape) == g[f'shape{i}_back'].shape
    assert_parity(cpu(back), g[f'shape{i}_back'], TOL, f'shape{i} back')


def test_leading_dims_compression_and_return_types():
    g = golden()
    x = randn((2, 3, 2500), 200)
    stft = brv.STFT(frame_length=510, hop_length=128, normalized=False,
                    compression_factor=0.5, scale_factor=0.15)
    spec = stft(x.to(DEV))
    assert tuple(spec.shape) == g['sgmse_spec'].shape == (2, 3, 256, 20)
    assert_parity(cpu(spec), g['sgmse_spec'], TOL)
    assert_parity(cpu(stft.backward(spec)), g['sgmse_back'], TOL)
    mag, phase = stft(x.to(DEV), return_type='mag_phase')
    assert_parity(cpu(mag), g['sgmse_mag'], TOL)
    assert_parity(cpu(torch.polar(mag, phase)), g['sgmse_spec'], TOL)
    re, im = stft(x.to(DEV), return_type='real_imag')
    assert_parity(cpu(torch.complex(re, im)), g['sgmse_spec'], TOL)
    assert_parity(cpu(stft.backward((re, im), input_type='real_imag')),
                  g['sgmse_back'], TOL)
    assert_parity(cpu(stft.backward((mag, phase), input_type='mag_phase')),
                  g['sgmse_back'], TOL)
    # TF-GridNet configuration: binaural in, one source out, frame-major view
    x = randn((3, 2, 3000), 201)
    stft = brv.STFT(frame_length=256, hop_length=128, normalized=False)
    spec = stft(x.to(DEV))
    assert_parity(cpu(spec), g['gridnet_spec'], TOL)
    assert_parity(cpu(stft.backward(spec[:, :1])), g['gridnet_back'], TOL)


def test_istft_layouts_and_input_preserved():
    """Bin-major contiguous (DCCRN/FFNN/SGMSE callers) and frame-major (our own
    output, TF-GridNet) inputs give the same result; the input is not modified."""
    g = golden()
    stft = brv.STFT(frame_length=512, hop_length=128, scale_factor=0.5)
    stft1 = brv.STFT(frame_length=512, hop_length=128)
    spec = crandn((2, 257, 20), 202).to(DEV)
    assert_parity(cpu(stft1.backward(spec)), g['istft_random'], TOL)
    keep = spec.clone()
    a = stft.backward(spec)                                   # bin-major
    b = stft.backward(spec.transpose(1, 2).contiguous().transpose(1, 2))
    assert torch.equal(spec, keep)
    assert_parity(cpu(a), cpu(b), 1e-6)
    assert_parity(cpu(a), O.istft(cpu(spec), 512, 128, scale_factor=0.5), TOL)


def test_empty_and_tiny_inputs():
    stft = brv.STFT()
    out = stft(torch.zeros(0, 1000, device=DEV))
    assert tuple(out.shape) == (0, 257, 5)
    out = stft(torch.ones(1, device=DEV))
    assert tuple(out.shape) == (257, 3)
    assert_parity(cpu(out), O.stft(np.ones(1)), TOL)
    y = stft.backward(out)
    assert tuple(y.shape) == (512,)
    assert abs(float(y[0]) - 1.0) < 1e-5


def test_nola_violation_raises_like_torch():
    window = np.zeros(512)
    window[:64] = 1.0
    stft = brv.STFT(window=window, hop_length=256)
    spec = stft(torch.randn(2048, device=DEV))
    with pytest.raises(RuntimeError, match='window overlap add min'):
        stft.backward(spec)


def test_nan_causality():
    """tests/test_models.py:57-80 style: a NaN at sample i only reaches frames
    whose support contains i (pins frame indexing / latency)."""
    stft = brv.STFT(512, 256)
    x = torch.randn(8192)
    i = 4000
    x[i] = float('nan')
    spec = stft(x.to(DEV))
    bad = torch.isnan(spec.real).any(0).cpu().numpy()
    t = np.arange(spec.shape[-1])
    expect = (t * 256 - 256 <= i) & (i < t * 256 - 256 + 512)
    assert np.array_equal(bad, expect)


def test_fp64_and_half_inputs():
    x = randn((3000,), 5)
    stft = brv.STFT()
    ref = O.stft(x.numpy())
    out64 = stft(x.double().to(DEV))
    assert out64.dtype == torch.complex128
    assert_parity(cpu(out64), ref, TOL)
    assert stft.backward(out64).dtype == torch.float64
    out16 = stft(x.half().to(DEV))
    assert out16.dtype == torch.complex64
    assert_parity(cpu(out16), O.stft(x.half().float().numpy()), TOL)


def test_stft_autograd_matches_reference_autograd():
    """Gradients of STFT.forward (needed by multiresyu) and STFT.backward (DCCRN,
    TF-GridNet) against torch autograd through the reference's library calls."""
    x = randn((2, 3000), 11)
    w = crandn((2, 257, 13), 12)
    win = torch.from_numpy(O.get_window('hann', 512))
    for kw in (dict(), dict(hop_length=128, normalized=False, scale_factor=0.3)):
        stft = brv.STFT(**kw)
        n_frames = stft.n_frames(3000)
        wt = crandn((2, 257, n_frames), 13)
        xr = x.clone().requires_grad_(True)
        (P.stft(xr, win, **kw) * wt.conj()).real.sum().backward()
        xg = x.clone().to(DEV).requires_grad_(True)
        (stft(xg) * wt.to(DEV).conj()).real.sum().backward()
        assert_parity(cpu(xg.grad), xr.grad.numpy(), TOL, 'd stft / dx')

        spec = crandn((2, 257, n_frames), 14)
        v = randn((2, stft.hop_length * (n_frames - 1)), 15)
        sr = spec.clone().requires_grad_(True)
        (P.istft(sr, win, **kw) * v).sum().backward()
        sg = spec.clone().to(DEV).requires_grad_(True)
        (stft.backward(sg) * v.to(DEV)).sum().backward()
        assert_parity(cpu(sg.grad), sr.grad.numpy(), TOL, 'd istft / dX')
        # bin-major AND frame-major gradient inputs
        sg2 = spec.transpose(1, 2).contiguous().transpose(1, 2).to(DEV).requires_grad_(True)
        (stft.backward(sg2) * v.to(DEV)).sum().backward()
        assert_parity(cpu(sg2.grad), sr.grad.numpy(), TOL)
    del w


# --------------------------------------------------------------------------- #
# full BASELINE sizes: properties + sampled oracle comparison                 #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize('name,shape,kw', [
    ('cfg1', (16, 2, 64000), dict(frame_length=512, hop_length=256)),
    ('cfg2', (64, 64000), dict(frame_length=512, hop_length=128)),
    ('cfg4', (16, 128000), dict(frame_length=510, hop_length=128,
                                normalized=False, compression_factor=0.5,
                                scale_factor=0.15)),
    ('cfg5', (64, 2, 64000), dict(frame_length=256, hop_length=128,
                                  normalized=False)),
])
def test_baseline_sizes_roundtrip_linearity_and_sampled_parity(name, shape, kw):
    mix, _ = synthetic_mixture(shape, 1000)
    stft = brv.STFT(**kw)
    x = mix.to(DEV)
    spec = stft(x)
    T = O.stft_frames(shape[-1], kw['frame_length'], kw['hop_length'])
    assert tuple(spec.shape) == (*shape[:-1], kw['frame_length'] // 2 + 1, T)
    # encode -> decode round trip (the reference's acceptance tolerances)
    y = stft.backward(spec)[..., :shape[-1]]
    assert torch.allclose(x, y, rtol=0, atol=1e-6 * float(x.abs().max()) * 10)
    assert_parity(cpu(y), mix.numpy(), 2e-5, name + ' roundtrip')
    # sampled value parity against the float64 oracle (first and last signals)
    flat = mix.reshape(-1, shape[-1])
    sflat = spec.reshape(-1, *spec.shape[-2:])
    for idx in (0, flat.shape[0] - 1):
        assert_parity(cpu(sflat[idx]), O.stft(flat[idx].numpy(), **kw), TOL,
                      f'{name} signal {idx}')
    if kw.get('compression_factor', 1) == 1:  # linearity (uncompressed only)
        a, b = x[:1], x[-1:]
        lin = stft(2 * a - 3 * b)
        assert_parity(cpu(lin), cpu(2 * stft(a) - 3 * stft(b)), 2e-5, name + ' linearity')


# --------------------------------------------------------------------------- #
# mel / features / FFNN glue                                                  #
# --------------------------------------------------------------------------- #
def test_mel_apply_forward_backward():
    g = golden()
    fb = brv.MelFilterbank()
    pw = randn((2, 257, 12), 300).abs()
    assert_parity(cpu(fb(pw.to(DEV))), g['mel_fwd'], TOL)
    assert_parity(cpu(fb.backward(randn((2, 64, 12), 301).to(DEV))), g['mel_bwd'], TOL)
    # unbatched + strided input
    assert_parity(cpu(fb(pw.to(DEV)[1])), g['mel_fwd'][1], TOL)
    tr = pw.transpose(1, 2).contiguous().transpose(1, 2).to(DEV)
    assert_parity(cpu(fb(tr)), g['mel_fwd'], TOL)
    # autograd = transposed filterbank
    xg = pw.clone().to(DEV).requires_grad_(True)
    fb(xg).sum().backward()
    ref = fb.filters.sum(0)[None, :, None].expand(2, 257, 12)
    assert_parity(cpu(xg.grad), ref.numpy(), TOL)


@pytest.mark.parametrize('name', ['fbe', 'logfbe', 'cubicfbe', 'pdf', 'logpdf',
                                  'cubicpdf'])
def test_feature_values(name):
    g = golden()
    fb = brv.MelFilterbank()
    fe = brv.FeatureExtractor([name], fb)
    su, sb = crandn((2, 257, 30), 400), crandn((4, 2, 257, 30), 401)
    out = fe(su.to(DEV))
    assert tuple(out.shape) == (64, 30)
    assert_parity(cpu(out), g[f'feat_u_{name}'], TOL, name)
    assert fe.indices == {name: (0, 64)}
    out = fe(sb.to(DEV))
    assert_parity(cpu(out), g[f'feat_b_{name}'], TOL, name)
    # frame-major input (the layout our STFT produces) gives the same values
    fm = sb.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2).to(DEV)
    assert_parity(cpu(fe(fm)), g[f'feat_b_{name}'], TOL, name + ' frame-major')


def test_feature_concat_quirk_and_errors():
    g = golden()
    fb = brv.MelFilterbank()
    fe = brv.FeatureExtractor(['logfbe', 'fbe', 'cubicpdf'], fb)
    su, sb = crandn((2, 257, 30), 400), crandn((4, 2, 257, 30), 401)
    out = fe(su.to(DEV))
    assert_parity(cpu(out), g['feat_multi_u'], TOL)
    assert [fe.indices[k] for k in sorted(fe.indices)] == \
        [tuple(r) for r in g['feat_multi_u_idx']]
    out = fe(sb.to(DEV))                       # dim-0 concat quirk (features.py:113)
    assert tuple(out.shape) == (12, 64, 30)
    assert_parity(cpu(out), g['feat_multi_b'], TOL)
    assert [fe.indices[k] for k in sorted(fe.indices)] == \
        [tuple(r) for r in g['feat_multi_b_idx']]
    with pytest.raises(ValueError):
        brv.FeatureExtractor(['nope'], fb)(su.to(DEV))


@pytest.mark.parametrize('name', ['ild', 'ipd', 'ic', 'mfcc', 'cubicmfcc', 'pdfcc'])
def test_binaural_and_dct_feature_values(name):
    """features.py:199-296: interaural level / phase difference, coherence (recursive spectra
    along frames, lfilter clamp included) and the DCT features with their deltas, against
    outputs of the reference itself and the float64 oracle."""
    g = golden()
    fb = brv.MelFilterbank()
    fe = brv.FeatureExtractor([name], fb)
    su, sb = crandn((2, 257, 30), 400), crandn((4, 2, 257, 30), 401)
    rows = 39 if name in ('mfcc', 'cubicmfcc', 'pdfcc') else 64
    out = fe(su.to(DEV))
    assert tuple(out.shape) == (rows, 30)
    assert_parity(cpu(out), g[f'feat_u_{name}'], TOL, name)
    assert fe.indices == {name: (0, rows)}
    assert_parity(cpu(fe(sb.to(DEV))), g[f'feat_b_{name}'], TOL, name)
    assert_parity(cpu(fe((0.1 * sb).to(DEV))), g[f'feat_bs_{name}'], TOL, name + ' 0.1x')
    # frame-major input (the layout our STFT produces) gives the same values
    fm = sb.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2).to(DEV)
    assert_parity(cpu(fe(fm)), g[f'feat_b_{name}'], TOL, name + ' frame-major')
    # tile seams of the frame-tiled kernel, long recursion: 700 frames against the oracle
    filters, _, _ = O.mel_filterbank()
    big = 0.3 * crandn((3, 2, 257, 700), 402)
    ref, _ = O.extract_features(big.numpy(), filters, [name])
    assert_parity(cpu(fe(big.to(DEV))), ref, TOL, name + ' 700 frames')


def test_mixed_features_and_ic_time_constant():
    g = golden()
    fb = brv.MelFilterbank()
    su, sb = crandn((2, 257, 30), 400), crandn((4, 2, 257, 30), 401)
    fe = brv.FeatureExtractor(['mfcc', 'ild', 'logfbe', 'ic', 'ipd'], fb)
    out = fe(su.to(DEV))
    assert_parity(cpu(out), g['feat_multi2_u'], TOL)
    assert [fe.indices[k] for k in sorted(fe.indices)] == \
        [tuple(r) for r in g['feat_multi2_u_idx']]
    assert fe.n_features == int(g['feat_multi2_n_features'][0])   # declared 13 vs returned 39
    fe = brv.FeatureExtractor(['ic'], fb, hop_length=64)
    assert_parity(cpu(fe((0.1 * sb).to(DEV))), g['feat_bs_ic_hop64'], TOL)
    with pytest.raises(IndexError):                 # mono input has no channel 1
        brv.FeatureExtractor(['ild'], fb)(su[:1].to(DEV))
    # features straight from our own STFT of a binaural mixture, against the oracle chain
    x = 0.05 * randn((2, 8000), 403)                # one unbatched binaural mixture
    stft = brv.STFT()
    fe = brv.FeatureExtractor(['ic', 'ild', 'ipd', 'pdfcc'], fb)
    got = fe(stft(x.to(DEV)))
    filters, _, _ = O.mel_filterbank()
    ref, _ = O.extract_features(O.stft(x.numpy()), filters, ['ic', 'ild', 'ipd', 'pdfcc'])
    assert got.shape == ref.shape == (64 * 3 + 39, 33)
    # (ipd of near-silent bins is ill-conditioned -- the angle of ~0 -- and is left out here)
    assert_parity(cpu(got[:64]), ref[:64], 2e-4, 'ic from STFT')
    assert_parity(cpu(got[64:128]), ref[64:128], 2e-4, 'ild from STFT')
    assert_parity(cpu(got[192:]), ref[192:], 2e-4, 'pdfcc from STFT')


def test_ffnn_glue_against_reference():
    g = golden()
    ffnn = brv.ffnn
    feats = randn((3, 64, 17), 501)
    assert np.array_equal(cpu(ffnn.stack(feats.to(DEV), 5)), g['ffnn_stack_b'])
    assert np.array_equal(cpu(ffnn.stack(feats[0].to(DEV), 5)), g['ffnn_stack_u'])
    stacked = ffnn.stack(feats.to(DEV), 5)
    norm = ffnn.StaticNormalizer(384).to(DEV)
    norm.set_statistics(torch.from_numpy(g['ffnn_static_mean']).to(DEV),
                        torch.from_numpy(g['ffnn_static_std']).to(DEV))
    assert_parity(cpu(norm(stacked)), g['ffnn_static'], TOL)
    cum = cpu(ffnn.CumulativeNormalizer()(stacked))
    assert_parity(cum, O.cumulative_normalize(cpu(stacked)), TOL)   # fp64 tie-breaker
    assert_parity(cum, g['ffnn_cumulative'], 3e-4)  # fp32 reference cancels (ffnn.py:199)
    # transform: STFT -> logfbe -> stack -> decimate ; irm labels
    sources = 0.05 * randn((2, 2, 4000), 500)
    front = ffnn.FFNNFrontEnd()
    out = front.transform(sources.to(DEV))
    assert tuple(out.shape) == g['ffnn_transform'].shape == (448, 17)
    assert_parity(cpu(out), g['ffnn_transform'], TOL)
    front3 = ffnn.FFNNFrontEnd(stacks=3, decimation=2)
    out = front3.transform(sources.to(DEV))
    assert tuple(out.shape) == g['ffnn_transform_s3d2'].shape
    assert_parity(cpu(out), g['ffnn_transform_s3d2'], TOL)
    # fused features + static normalisation == separate calls
    spec = front.stft(sources.to(DEV))
    mean, std = randn((384, 1), 502), randn((384, 1), 503).abs() + 0.5
    fused = front.features(spec, mean, std)
    sep = (front.features(spec) - mean.to(DEV)) / std.to(DEV)
    assert_parity(cpu(fused), cpu(sep), 1e-6)
    # enhance tail with a fixed mask in place of the MLP (ffnn.py:105-110)
    mix = 0.05 * randn((3, 2, 4000), 504)
    mask = torch.sigmoid(randn((3, 64, 17), 505)).to(DEV)
    y = front.enhance(mix.to(DEV), lambda feats: mask)
    assert tuple(y.shape) == (3, 4000)
    assert_parity(cpu(y), g['ffnn_enh_out'], TOL)
    assert_parity(cpu(front.mel_fb.backward(mask)), g['ffnn_enh_mask_ext'], TOL)


def test_ffnn_front_end_with_binaural_features():
    """FFNN with features={'ild', 'ic', 'logfbe'} (config/models/ffnn.yaml allows any subset):
    per-feature kernels, concatenated on the feature axis, stacked and decimated."""
    mix, _ = synthetic_mixture((3, 2, 16000), 1002)
    front = brv.ffnn.FFNNFrontEnd(features=('ild', 'ic', 'logfbe'), stacks=2, decimation=2)
    assert front.input_size == 3 * 64 * 3
    feats = front.features(front.stft(mix.to(DEV)))
    filters, _, _ = O.mel_filterbank()
    for idx in range(3):
        spec = O.stft(mix[idx].numpy())
        parts = [O.ic(spec, filters), O.ild(spec, filters), O.fbe(spec, filters, compression='log')]
        ref = O.decimate(O.stack(np.concatenate(parts, axis=0), 2), 2)
        assert tuple(feats[idx].shape) == ref.shape
        assert_parity(cpu(feats[idx]), ref, 2e-4, f'item {idx}')


def test_features_baseline_size_against_oracle():
    """cfg1: 16 x 2ch x 4 s, 512/256, 64 log-mel, 5 stacks + static stats."""
    mix, _ = synthetic_mixture((16, 2, 64000), 1001)
    front = brv.ffnn.FFNNFrontEnd()
    spec = front.stft(mix.to(DEV))
    feats = front.features(spec)
    assert tuple(feats.shape) == (16, 384, 251)
    filters, _, _ = O.mel_filterbank()
    for idx in (0, 15):
        ref = O.stack(O.fbe(O.stft(mix[idx].numpy()), filters, compression='log'), 5)
        assert_parity(cpu(feats[idx]), ref, TOL, f'item {idx}')
    mean, std = brv.ffnn.training_statistics([f for f in feats])
    normed = front.features(spec, mean, std)
    assert_parity(cpu(normed), cpu((feats - mean) / std), 1e-5)
    assert abs(float(normed.mean())) < 1e-3


def test_conv_stft_against_reference():
    """ConvSTFT (stft.py:201-319) on the folded tensor-core kernels: the reference's own test
    input and parametrisation (tests/test_modules.py:329-352) against outputs of the reference,
    its round-trip acceptance property, ragged / batched shapes, attributes."""
    g = golden()
    x = randn((4096,), 42)
    lib = _lib.lib()
    for hop, c, s, n in itertools.product([256, 128], [1.0, 0.5], [1.0, 0.15], [False, True]):
        kw = dict(frame_length=512, hop_length=hop, compression_factor=c, scale_factor=s,
                  normalized=n)
        key = f'conv_h{hop}_c{c}_s{s}_n{int(n)}'
        cs = brv.ConvSTFT(**kw)
        n0 = lib.brv_launch_count()
        spec = cs(x.to(DEV))
        assert lib.brv_launch_count() - n0 == 1
        ref = O.conv_stft(x.numpy(), **kw)
        assert tuple(spec.shape) == ref.shape
        if key + '_spec' in g.files:
            assert_parity(cpu(spec), g[key + '_spec'], TOL, key)
        assert_parity(cpu(spec), ref, TOL, key + ' oracle')
        back = cs.backward(spec)
        assert_parity(cpu(back), g[key + '_back'], TOL, key + ' back')
        assert torch.allclose(x, back.cpu(), rtol=1e-1, atol=1e-1)   # tests/test_modules.py:352
    for i in range(5):
        S, L, H = (int(v) for v in g[f'convshape{i}_meta'])
        xs = randn((2, 3, S), 150 + i)
        cs = brv.ConvSTFT(frame_length=L, hop_length=H)
        spec = cs(xs.to(DEV))
        assert tuple(spec.shape) == g[f'convshape{i}_spec'].shape
        assert_parity(cpu(spec), g[f'convshape{i}_spec'], TOL, f'shape {i}')
        back = cs.backward(spec)
        assert tuple(back.shape) == g[f'convshape{i}_back'].shape   # (…, 0) when hop == L
        assert_parity(cpu(back), g[f'convshape{i}_back'], TOL, f'shape {i} back')
        re, im = cs(xs.to(DEV), return_type='real_imag')
        assert_parity(cpu(cs.backward((re, im), input_type='real_imag')),
                      g[f'convshape{i}_back'], TOL)
    cs = brv.ConvSTFT(frame_length=512, hop_length=128)
    assert_parity(cpu(cs.backward(crandn((2, 257, 20), 160).to(DEV))), g['conv_random_back'], TOL)
    # bin-major and frame-major inputs, a long signal crossing several 128-frame tiles
    xs = randn((3, 40000), 161)
    spec = cs(xs.to(DEV))
    ref = O.conv_stft(xs.numpy(), 512, 128)
    assert_parity(cpu(spec), ref, TOL)
    ref_back = O.conv_istft(ref, 512, 128)
    assert_parity(cpu(cs.backward(spec)), ref_back, TOL)
    assert_parity(cpu(cs.backward(spec.contiguous())), ref_back, TOL)
    # attributes the reference exposes
    assert tuple(cs.filters.shape) == (514, 1, 512) and cs.filters.dtype == torch.float32
    assert cs._normalization_factor == 0.5 * 512 / 128 ** 0.5
    assert cs.frame_count(4096) == 29 and tuple(cs.pad(xs).shape) == (3, 40064 + 768)
    with pytest.raises(ValueError):
        cs(xs.to(DEV), return_type='nope')
    # sizes outside the folded tensor-core kernels run on the direct-sum kernels (round 2)
    c2 = brv.ConvSTFT(frame_length=400, hop_length=100)
    s2 = c2(xs.to(DEV))
    assert_parity(cpu(s2), O.conv_stft(xs.numpy(), 400, 100), TOL)


# --------------------------------------------------------------------------- #
# criteria                                                                    #
# --------------------------------------------------------------------------- #
def _crit_inputs():
    B, S, L = 5, 3, 2000
    est, ref = randn((B, S, L), 600), randn((B, S, L), 601)
    est = ref.roll(1, 1) * 0.7 + 0.3 * est
    return est, ref, torch.tensor(golden()['crit_lengths'])


def test_criteria_values_against_reference():
    g = golden()
    est, ref, lengths = _crit_inputs()
    e, r = est.to(DEV), ref.to(DEV)
    out = brv.snr(e, r, lengths)
    assert tuple(out.shape) == (5,) and out.dtype == torch.float32
    assert np.allclose(cpu(out), g['crit_snr'], rtol=0, atol=2e-5)
    assert np.allclose(cpu(out), g['crit_snr_f64'], rtol=0, atol=2e-6)
    out = brv.sisnr(e, r, lengths.to(DEV))
    assert np.allclose(cpu(out), g['crit_sisnr'], rtol=0, atol=2e-5)
    assert np.allclose(cpu(out), g['crit_sisnr_f64'], rtol=0, atol=2e-6)
    # 2-D input: torch's mean(()) quirk -> 0-dim batch mean (DCCRN path)
    out = brv.snr(e[:, 0], r[:, 0], lengths)
    assert out.ndim == 0
    assert abs(float(out) - float(g['crit_snr_2d'])) < 2e-5
    # 4-D input, non-contiguous rows (sliced last dim)
    e4 = e.view(5, 3, 2, 1000)[..., :900]
    r4 = r.view(5, 3, 2, 1000)[..., :900]
    out = brv.snr(e4, r4, lengths.clamp(max=900))
    assert np.allclose(cpu(out), g['crit_snr_4d'], rtol=0, atol=2e-5)
    # high SNR: the float64 oracle is the tie-breaker (fp32 reference loses digits)
    close = (ref + 1e-4 * randn((5, 3, 2000), 602))
    n = lengths.numpy()
    assert np.allclose(cpu(brv.snr(close.to(DEV), r, lengths)),
                       O.snr(close.numpy(), ref.numpy(), n), atol=1e-4)
    assert np.allclose(cpu(brv.sisnr(close.to(DEV), r, lengths)),
                       O.sisnr(close.numpy(), ref.numpy(), n)[0], atol=1e-4)
    assert np.allclose(cpu(brv.snr(close.to(DEV), r, lengths)), g['crit_snr_close'], atol=2e-2)
    # float64 in -> float64 out
    assert brv.snr(e.double(), r.double(), lengths).dtype == torch.float64
    # metrics are -criterion (metrics.py:112-123)
    assert float(brv.snr(r, r, lengths).max()) < -60


def test_criteria_gradients():
    g = golden()
    est, ref, lengths = _crit_inputs()
    e = est.clone().to(DEV).requires_grad_(True)
    brv.snr(e, ref.to(DEV), lengths).sum().backward()
    assert_parity(cpu(e.grad), g['crit_snr_grad'], TOL)
    assert float(e.grad[1, :, 1500:].abs().max()) == 0     # masked tail
    e = est.clone().to(DEV).requires_grad_(True)
    brv.sisnr(e, ref.to(DEV), lengths).sum().backward()
    assert_parity(cpu(e.grad), g['crit_sisnr_grad'], TOL)
    # 2-D quirk gradient: d mean_b / dx
    e = est[:, 0].clone().to(DEV).requires_grad_(True)
    brv.snr(e, ref[:, 0].to(DEV), lengths).backward()
    er = est[:, 0].clone().requires_grad_(True)
    P.snr(er, ref[:, 0], lengths).backward()
    assert_parity(cpu(e.grad), er.grad.numpy(), TOL)
    # weighted upstream gradient
    wgt = torch.tensor([1.0, -2.0, 0.5, 3.0, 0.0])
    e = est.clone().to(DEV).requires_grad_(True)
    (brv.sisnr(e, ref.to(DEV), lengths) * wgt.to(DEV)).sum().backward()
    er = est.clone().requires_grad_(True)
    (P.sisnr(er, ref, lengths) * wgt).sum().backward()
    assert_parity(cpu(e.grad), er.grad.numpy(), TOL)


@pytest.mark.parametrize('name', ['snr', 'sisnr'])
def test_batched_equals_single_reference_sizes(name):
    """tests/test_losses.py:13-57, same sizes (B=16, S=4, 16000..32000)."""
    torch.manual_seed(0)
    B, S, lo, hi = 16, 4, 16000, 32000
    lengths = torch.randint(lo, hi, (B,))
    inputs = [torch.randn(S, n) for n in lengths]
    batched_in = torch.stack([torch.nn.functional.pad(x, (0, hi - x.shape[-1]))
                              for x in inputs])
    batched_out = batched_in + torch.randn(*batched_in.shape)   # padding NOT zero
    targets = [torch.randn(S, n) for n in lengths]
    batched_tgt = torch.stack([torch.nn.functional.pad(x, (0, hi - x.shape[-1]))
                               for x in targets])
    crit = brv.init_criterion(name)
    batched = crit(batched_out.to(DEV), batched_tgt.to(DEV), lengths)
    single = torch.stack([
        crit(batched_out[i:i + 1, :, :n].to(DEV), targets[i][None].to(DEV),
             torch.tensor([n]))[0] for i, n in enumerate(lengths)])
    assert torch.allclose(batched, single)
    ref = (P.snr if name == 'snr' else P.sisnr)(batched_out, batched_tgt, lengths)
    assert np.allclose(cpu(batched), ref.numpy(), rtol=0, atol=5e-5)


def test_mse_against_reference():
    """criterion.py:104-132: values, weights, the 2-D mean(()) quirk, complex inputs, gradients."""
    g = golden()
    est, ref, lengths = _crit_inputs()
    weight = torch.tensor(g['crit_mse_weight'])
    e, r = est.to(DEV), ref.to(DEV)
    assert np.allclose(cpu(brv.mse(e, r, lengths)), g['crit_mse'], rtol=2e-6)
    assert np.allclose(cpu(brv.mse(e, r, lengths, weight=weight.to(DEV))), g['crit_mse_w'], rtol=2e-6)
    out = brv.mse(e[:, 0], r[:, 0], lengths)
    assert out.ndim == 0 and abs(float(out) - float(g['crit_mse_2d'])) < 1e-5
    cest = torch.complex(est[..., :1000], est[..., 1000:])
    cref = torch.complex(ref[..., :1000], ref[..., 1000:])
    clen = lengths.clamp(max=1000)
    out = brv.mse(cest.to(DEV), cref.to(DEV), clen, weight=weight.to(DEV))
    assert out.dtype == torch.float32
    assert np.allclose(cpu(out), g['crit_mse_complex'], rtol=2e-6)
    eg = est.clone().to(DEV).requires_grad_(True)
    brv.mse(eg, r, lengths, weight=weight.to(DEV)).sum().backward()
    assert_parity(cpu(eg.grad), g['crit_mse_grad'], 1e-5)
    assert float(eg.grad[3, :, 801:].abs().max()) == 0           # masked tail
    cg = cest.clone().to(DEV).requires_grad_(True)
    brv.mse(cg, cref.to(DEV), clen).sum().backward()
    assert_parity(cpu(cg.grad), g['crit_mse_complex_grad'], 1e-5)
    assert brv.init_criterion('mse') is brv.mse


MRY_KW = {'def': {}, 'multi': dict(frame_lengths=[512, 256], hop_lengths=[128, 128],
                                   time_domain_weight=0.3, spectral_weight=0.7),
          'si': dict(scale_invariant=True)}


@pytest.mark.parametrize('name', ['def', 'multi', 'si'])
def test_multiresyu_against_reference(name):
    """criterion.py:135-226 (TF-GridNet's default criterion): values, 2-D quirk, gradients
    through the fused L1 reductions and the tcgen05 STFT gradient kernel."""
    g = golden()
    est, ref, lengths = _crit_inputs()
    crit = brv.init_criterion('multiresyu', **MRY_KW[name])
    e, r = est.to(DEV), ref.to(DEV)
    assert np.allclose(cpu(crit(e, r, lengths)), g[f'crit_mry_{name}'], rtol=2e-5)
    out = crit(e[:, 0], r[:, 0], lengths)
    assert out.ndim == 0 and abs(float(out) - float(g[f'crit_mry_{name}_2d'])) < 2e-4
    eg = est.clone().to(DEV).requires_grad_(True)
    weight = torch.tensor(g['crit_mse_weight']).to(DEV)
    (crit(eg, r, lengths) * weight).sum().backward()
    assert_parity(cpu(eg.grad), g[f'crit_mry_{name}_grad'], 1e-4)
    assert float(eg.grad[3, :, 801:].abs().max()) == 0


@pytest.mark.parametrize('name', ['mse', 'multiresyu'])
def test_batched_equals_single_next_criteria(name):
    """tests/test_losses.py:13-57 for the two remaining registry entries."""
    torch.manual_seed(0)
    B, S, lo, hi = 8, 4, 16000, 32000
    lengths = torch.randint(lo, hi, (B,))
    batched_out = torch.randn(B, S, hi)                          # padding NOT zero
    batched_tgt = torch.randn(B, S, hi)
    crit = brv.init_criterion(name)
    batched = crit(batched_out.to(DEV), batched_tgt.to(DEV), lengths)
    single = torch.stack([
        crit(batched_out[i:i + 1, :, :n].to(DEV), batched_tgt[i:i + 1, :, :n].to(DEV),
             torch.tensor([n]))[0] for i, n in enumerate(lengths)])
    assert torch.allclose(batched, single, rtol=1e-5)
    port = P.mse if name == 'mse' else P.multiresyu
    assert np.allclose(cpu(batched), port(batched_out, batched_tgt, lengths).numpy(), rtol=2e-5)


def test_sisnr_baseline_size_and_pit():
    """cfg3 shape: (256, 1, 64000) and the PIT S=2 variant with swapped sources."""
    mix, fg = synthetic_mixture((256, 1, 64000), 1003)
    lengths = torch.full((256,), 64000)
    out = brv.sisnr(mix.to(DEV), fg.to(DEV), lengths)
    ref, _ = O.sisnr(mix[:4].numpy(), fg[:4].numpy(), lengths[:4].numpy())
    assert np.allclose(cpu(out[:4]), ref, atol=1e-4)
    assert np.allclose(cpu(out[:4]), P.sisnr(mix[:4], fg[:4], lengths[:4]).numpy(), atol=1e-3)
    two = torch.cat([fg[:32], mix[:32] - fg[:32]], 1)            # (32, 2, L)
    est = two.flip(1) + 0.01 * torch.randn_like(two)
    swapped = brv.sisnr(est.to(DEV), two.to(DEV), lengths[:32])
    direct = brv.sisnr(est.flip(1).to(DEV), two.to(DEV), lengths[:32])
    assert torch.allclose(swapped, direct, atol=1e-4)            # PIT finds the swap
    assert float(swapped.max()) < -5


def test_apply_mask():
    x, y = randn((3, 2, 50), 1), randn((3, 2, 50), 2)
    lengths = torch.tensor([50, 0, 17])
    mx, my = brv.apply_mask(x.to(DEV), y.to(DEV), lengths)
    rx, ry = O.apply_mask(x.numpy(), y.numpy(), lengths.numpy())
    assert np.array_equal(cpu(mx), rx.astype(np.float32))
    assert np.array_equal(cpu(my), ry.astype(np.float32))


def test_criterion_workspace_reuse_across_layouts():
    """A few-pairs / long-rows call leaves partial sums where a later many-pairs call puts its
    ticket counters (cached workspace): the counters must be re-zeroed, or the last-CTA
    finalise never runs and the loss is garbage."""
    from brever_b200 import criterion as C
    C._workspaces.clear()
    for crit, port in ((brv.snr, P.snr), (brv.sisnr, P.sisnr)):
        a_x, a_y = randn((8, 1, 163840), 11), randn((8, 1, 163840), 12)
        la = torch.full((8,), 163840)
        first = crit(a_x.to(DEV), a_y.to(DEV), la)
        assert np.allclose(cpu(first), port(a_x, a_y, la).numpy(), atol=1e-3)
        b_x, b_y = randn((70, 1, 16000), 13), randn((70, 1, 16000), 14)
        lb = torch.full((70,), 16000)
        second = crit(b_x.to(DEV), b_y.to(DEV), lb)
        assert np.allclose(cpu(second), port(b_x, b_y, lb).numpy(), atol=1e-3)
        # and back again, then a PIT call with S^2 pairs per item
        again = crit(a_x.to(DEV), a_y.to(DEV), la)
        assert torch.equal(first, again)
    two_x, two_y = randn((32, 2, 8000), 15), randn((32, 2, 8000), 16)
    lt = torch.full((32,), 8000)
    pit = brv.sisnr(two_x.to(DEV), two_y.to(DEV), lt)
    assert np.allclose(cpu(pit), P.sisnr(two_x, two_y, lt).numpy(), atol=1e-3)


def test_return_and_input_types_fused():
    """stft.py:91-110 `return_type` / `input_type` and MetricGAN-OKD's log1p / expm1 wrappers
    (metricganokd.py:185-195) through brv_spec_split / brv_spec_join: values, strides and
    gradients against the eager torch ops the reference runs."""
    from brever_b200.modules import specfmt
    stft = brv.STFT(512, 128)
    x = randn((2, 3, 3000), 5).to(DEV)
    spec = stft(x)
    re, im = stft(x, return_type='real_imag')
    assert torch.equal(re, spec.real) and torch.equal(im, spec.imag)
    mag, ph = stft(x, return_type='mag_phase')
    assert mag.shape == spec.shape and mag.stride() == spec.abs().stride()
    assert torch.allclose(mag, spec.abs(), rtol=2e-6, atol=0)
    assert torch.allclose(ph, spec.angle(), rtol=0, atol=2e-6)
    y = stft.backward(spec)
    assert torch.allclose(stft.backward((mag, ph), input_type='mag_phase'), y, atol=2e-6)
    assert torch.equal(stft.backward((re, im), input_type='real_imag'), y)
    # planes in the layout a network produces (bin-major contiguous) go through as well
    assert torch.allclose(stft.backward((mag.contiguous(), ph.contiguous()), input_type='mag_phase'), y, atol=2e-6)
    eps = 1e-7
    a, b = specfmt.split(spec, 'log1p_mag_phase', eps)
    assert torch.allclose(a, torch.log1p(spec.abs() + eps), rtol=2e-6, atol=1e-7)
    back = specfmt.join(a, b, 'log1p_mag_phase')
    ref = torch.expm1(a) * torch.exp(1j * b)
    assert rel_err(cpu(back), cpu(ref))[0] < 2e-6
    with pytest.raises(ValueError):
        specfmt.split(spec, 'polar')
    # gradients against torch autograd through the reference's eager ops
    for kind in ('real_imag', 'mag_phase', 'log1p_mag_phase'):
        X = crandn((2, 40, 30), 3).to(DEV)
        w1, w2 = randn((2, 40, 30), 4).to(DEV), randn((2, 40, 30), 6).to(DEV)

        def ref_split(X):
            if kind == 'real_imag':
                return X.real, X.imag
            m = X.abs()
            return (torch.log1p(m + eps) if kind == 'log1p_mag_phase' else m), X.angle()
        Xa, Xb = X.clone().requires_grad_(True), X.clone().requires_grad_(True)
        pa = specfmt.split(Xa, kind, eps)
        (pa[0] * w1 + pa[1] * w2).sum().backward()
        pb = ref_split(Xb)
        (pb[0] * w1 + pb[1] * w2).sum().backward()
        assert rel_err(cpu(Xa.grad), cpu(Xb.grad))[0] < 1e-5, kind
        m0, p0 = randn((2, 40, 30), 8).abs().to(DEV) + 0.1, randn((2, 40, 30), 9).to(DEV)
        W = crandn((2, 40, 30), 10).to(DEV)
        outs = []
        for fn in ('ours', 'torch'):
            m, p = m0.clone().requires_grad_(True), p0.clone().requires_grad_(True)
            if fn == 'ours':
                Z = specfmt.join(m, p, kind)
            elif kind == 'real_imag':
                Z = torch.complex(m, p)
            else:
                Z = torch.polar(torch.expm1(m) if kind == 'log1p_mag_phase' else m, p)
            (Z * W).real.sum().backward()
            outs.append((m.grad, p.grad))
        assert rel_err(cpu(outs[0][0]), cpu(outs[1][0]))[0] < 1e-5, kind
        assert rel_err(cpu(outs[0][1]), cpu(outs[1][1]))[0] < 1e-5, kind


def test_lazy_conjugate_inputs_are_resolved():
    """A lazily conjugated spectrogram (conj bit set, memory un-conjugated) must be read as its
    conjugate: raw data pointers go to the kernels."""
    stft = brv.STFT(256, 64)
    spec = crandn((2, 129, 50), 12).to(DEV)
    lazy = spec.conj()
    assert lazy.is_conj()
    assert torch.equal(stft.backward(lazy), stft.backward(lazy.resolve_conj()))
    fe = brv.FeatureExtractor(features=['logfbe'], mel_fb=brv.MelFilterbank(n_fft=256))
    four = crandn((1, 2, 129, 20), 13).to(DEV)
    assert torch.equal(fe(four.conj()), fe(four.conj().resolve_conj()))


def test_metric_wrappers():
    """brever/metrics.py:112-123: negated criteria, default lengths, .item() for 1-D input."""
    x, y = randn((4, 3000), 21), randn((4, 3000), 22)
    lengths = torch.tensor([3000, 2500, 100, 3000])
    for name, port in (('snr', P.snr), ('sisnr', P.sisnr)):
        fn = brv.metrics.MetricRegistry.get(name)
        out = fn(x.to(DEV), y.to(DEV), lengths)
        ref = -port(x[:, None], y[:, None], lengths)
        assert out.shape == (4,) and np.allclose(cpu(out), ref.numpy(), atol=5e-5)
        one = fn(x[0].to(DEV), y[0].to(DEV))
        assert isinstance(one, float) and abs(one - float(ref[0])) < 5e-5
