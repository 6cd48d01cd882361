"""Pin the CPU oracle (numpy restatement + torch port) against vectors produced
by the reference itself (tests/golden/make_golden.py) and against the
properties the reference's own tests assert."""
import itertools

import numpy as np
import pytest
import torch

from oracle import tf_oracle as O
from oracle import torch_port as P

from _util import (STFT_SHAPE_CASES, assert_parity, crandn, golden, randn,
                   rel_err)

COMBOS = list(itertools.product([256, 128], [1.0, 0.5], [1.0, 0.15],
                                [False, True], [False, True]))


@pytest.mark.parametrize('hop,c,s,normalized,onesided', COMBOS)
def test_reference_roundtrip_cases(hop, c, s, normalized, onesided):
    """tests/test_modules.py:300-326 input; spectra + round trip vs golden."""
    g = golden()
    x = randn((4096,), 42).numpy()
    kw = dict(frame_length=512, hop_length=hop, compression_factor=c,
              scale_factor=s, normalized=normalized, onesided=onesided)
    key = f'rt_h{hop}_c{c}_s{s}_n{int(normalized)}_o{int(onesided)}'
    spec = O.stft(x, **kw)
    if key + '_spec' in g:
        assert_parity(spec, g[key + '_spec'], 2e-6, key)
    back = O.istft(spec, **kw)
    assert back.shape == g[key + '_back'].shape
    assert_parity(back, g[key + '_back'], 5e-6, key + ' back')
    # the reference's own acceptance test, applied to the oracle
    assert np.allclose(x, back, rtol=0, atol=1e-6)
    assert np.allclose(x, back, rtol=2e-3, atol=0)
    # torch port agrees too
    win = torch.from_numpy(O.get_window('hann', 512))
    tp = P.stft(torch.from_numpy(x), win, **kw)
    if key + '_spec' in g:
        assert_parity(tp.numpy(), g[key + '_spec'], 1e-6, key + ' port')
    assert_parity(P.istft(tp, win, **kw).numpy(), g[key + '_back'], 1e-5,
                  key + ' port back')


@pytest.mark.parametrize('i', range(len(STFT_SHAPE_CASES)))
def test_shape_cases(i):
    g = golden()
    S, L, H, nfft, win = STFT_SHAPE_CASES[i]
    meta = g[f'shape{i}_meta']
    assert tuple(meta) == (S, L, H, nfft or L)
    x = randn((2, S), 100 + i).numpy()
    ref = g[f'shape{i}_spec']
    # integer frame arithmetic: bit exact
    T = O.stft_frames(S, L, H, nfft)
    assert ref.shape == (2, (nfft or L) // 2 + 1, T)
    assert g[f'shape{i}_back'].shape == (2, O.istft_length(T, H, nfft or L))
    spec = O.stft(x, frame_length=L, hop_length=H, window=win, n_fft=nfft)
    assert_parity(spec, ref, 2e-6, f'shape{i}')
    back = O.istft(ref, frame_length=L, hop_length=H, window=win, n_fft=nfft)
    assert_parity(back, g[f'shape{i}_back'], 5e-6, f'shape{i} back')


def test_frame_arithmetic_known_values():
    """SURVEY §8 probe values, bit-checked against the reference there."""
    assert O.stft_frames(64000, 512, 256) == 251
    assert O.stft_frames(64000, 512, 128) == 501
    assert O.stft_frames(128000, 510, 128) == 1001
    assert O.right_padding(128000, 510, 128) == 126
    assert O.stft_frames(64000, 256, 128) == 501
    assert O.right_padding(63999, 512, 256) == 1
    assert O.stft_frames(63999, 512, 256) == 251
    assert O.right_padding(100, 512, 256) == 412
    assert O.stft_frames(100, 512, 256) == 3
    assert O.stft_frames(4000, 512, 256) == 17
    assert O.istft_length(17, 256, 512) == 4096
    assert O.stft_frames(16000, 400, 100, 512) == 161


def test_windows_match_scipy():
    import scipy.signal
    for name in ['hann', 'hamming', 'blackman', 'boxcar']:
        for n in [512, 510, 256, 400]:
            assert np.allclose(O.get_window(name, n),
                               scipy.signal.get_window(name, n), atol=1e-15)
    assert abs((O.get_window('hann', 512) ** 2).sum() - 192.0) < 1e-12


def test_sgmse_and_gridnet_configs():
    g = golden()
    x = randn((2, 3, 2500), 200).numpy()
    kw = dict(frame_length=510, hop_length=128, normalized=False,
              compression_factor=0.5, scale_factor=0.15)
    spec = O.stft(x, **kw)
    # |X|^0.5 amplifies the float32 reference's own rounding near small bins
    assert_parity(spec, g['sgmse_spec'], 1e-5)
    assert_parity(O.istft(g['sgmse_spec'], **kw), g['sgmse_back'], 1e-5)
    assert_parity(np.abs(spec), g['sgmse_mag'], 1e-5)
    x = randn((3, 2, 3000), 201).numpy()
    kw = dict(frame_length=256, hop_length=128, normalized=False)
    assert_parity(O.stft(x, **kw), g['gridnet_spec'], 2e-6)
    assert_parity(O.istft(g['gridnet_spec'][:, :1], **kw), g['gridnet_back'],
                  5e-6)
    spec = crandn((2, 257, 20), 202).numpy()
    assert_parity(O.istft(spec, 512, 128), g['istft_random'], 5e-6)


@pytest.mark.parametrize('tag,kw', [
    ('mel512', {}), ('mel256', dict(n_fft=256)),
    ('mel40', dict(n_filters=40, n_fft=400, fs=8000, fmax=4000))])
def test_mel_constants(tag, kw):
    """Same support (non-zero pattern) as the reference, values within 1 ulp
    of float32 (torch's vectorised powf is not reproducible bit-for-bit in
    numpy; see oracle/tf_oracle.py:mel_filterbank)."""
    g = golden()
    filters, fc, scaling = O.mel_filterbank(**kw)
    assert np.allclose(fc, g[tag + '_fc'], rtol=2e-6, atol=0)
    assert np.array_equal(filters != 0, g[tag + '_filters'] != 0)
    assert np.allclose(filters, g[tag + '_filters'], rtol=0, atol=1e-5)
    assert np.allclose(scaling, g[tag + '_scaling'], rtol=1e-5, atol=0)
    assert np.allclose((filters * scaling).T, g[tag + '_inverse'], atol=2e-5)


def test_mel_structure():
    """SURVEY §8 a5 probe: 497 non-zeros, <= 2 filters per bin, rows sum to 1."""
    filters, _, _ = O.mel_filterbank()
    assert (filters != 0).sum() == 497
    assert ((filters != 0).sum(0) <= 2).all()
    assert np.allclose(filters.sum(1), 1, atol=1e-6)
    assert not filters[:, [0, 1, 256]].any()


def test_mel_apply():
    g = golden()
    filters, _, scaling = O.mel_filterbank()
    pw = randn((2, 257, 12), 300).abs().numpy()
    assert_parity(O.mel_forward(filters, pw), g['mel_fwd'], 2e-6)
    assert_parity(O.mel_backward(filters, scaling,
                                 randn((2, 64, 12), 301).numpy()),
                  g['mel_bwd'], 2e-6)


def test_conv_stft_against_reference():
    """ConvSTFT (stft.py:201-319): the reference's own test input and all 16 parametrisations
    (tests/test_modules.py:329-352), its acceptance property, and ragged / batched shapes
    including the hop == frame_length case where the reference returns nothing."""
    import itertools
    g = golden()
    x = randn((4096,), 42).numpy()
    for hop, c, s, n in itertools.product([256, 128], [1.0, 0.5], [1.0, 0.15], [False, True]):
        kw = dict(frame_length=512, hop_length=hop, compression_factor=c, scale_factor=s,
                  normalized=n)
        key = f'conv_h{hop}_c{c}_s{s}_n{int(n)}'
        spec = O.conv_stft(x, **kw)
        if key + '_spec' in g.files:
            assert_parity(spec, g[key + '_spec'], 5e-6, key)
        back = O.conv_istft(spec, **kw)
        assert_parity(back, g[key + '_back'], 5e-6, key)
        assert np.allclose(x, back, rtol=1e-1, atol=1e-1)     # tests/test_modules.py:352
    for i in range(5):
        S, L, H = (int(v) for v in g[f'convshape{i}_meta'])
        xs = randn((2, 3, S), 150 + i).numpy()
        spec = O.conv_stft(xs, frame_length=L, hop_length=H)
        assert spec.shape == g[f'convshape{i}_spec'].shape
        assert_parity(spec, g[f'convshape{i}_spec'], 5e-6, f'shape {i}')
        back = O.conv_istft(spec, frame_length=L, hop_length=H)
        assert back.shape == g[f'convshape{i}_back'].shape
        assert_parity(back, g[f'convshape{i}_back'], 5e-6, f'shape {i}')
    assert_parity(O.conv_istft(crandn((2, 257, 20), 160).numpy(), 512, 128),
                  g['conv_random_back'], 5e-6)


@pytest.mark.parametrize('name', sorted(O.FBE_FAMILY))
def test_features(name):
    g = golden()
    filters, _, _ = O.mel_filterbank()
    su, sb = crandn((2, 257, 30), 400).numpy(), crandn((4, 2, 257, 30), 401).numpy()
    assert_parity(O.fbe(su, filters, **O.FBE_FAMILY[name]),
                  g[f'feat_u_{name}'], 5e-6, name)
    assert_parity(O.fbe(sb, filters, **O.FBE_FAMILY[name]),
                  g[f'feat_b_{name}'], 5e-6, name)


@pytest.mark.parametrize('name', ['ild', 'ipd', 'ic', 'mfcc', 'cubicmfcc', 'pdfcc'])
def test_binaural_and_dct_features(name):
    """features.py:199-296 against outputs of the reference itself (unbatched, batched, and a
    0.1x batch that keeps ic's recursive spectra below torchaudio.lfilter's clamp)."""
    g = golden()
    filters, _, _ = O.mel_filterbank()
    su, sb = crandn((2, 257, 30), 400).numpy(), crandn((4, 2, 257, 30), 401).numpy()
    for tag, spec in (('u', su), ('b', sb), ('bs', 0.1 * sb)):
        out, idx = O.extract_features(spec, filters, [name])
        assert_parity(out, g[f'feat_{tag}_{name}'], 5e-6, f'{name} {tag}')
    rows = 39 if name in O.DCT_FAMILY else 64
    assert out.shape == (4, rows, 30)


def test_ic_time_constant_and_mixed_features():
    g = golden()
    filters, _, _ = O.mel_filterbank()
    su, sb = crandn((2, 257, 30), 400).numpy(), crandn((4, 2, 257, 30), 401).numpy()
    out, _ = O.extract_features(0.1 * sb, filters, ['ic'], hop_length=64)
    assert_parity(out, g['feat_bs_ic_hop64'], 5e-6)
    out, idx = O.extract_features(su, filters, ['mfcc', 'ild', 'logfbe', 'ic', 'ipd'])
    assert_parity(out, g['feat_multi2_u'], 5e-6)
    assert [idx[k] for k in sorted(idx)] == [tuple(r) for r in g['feat_multi2_u_idx']]
    # the reference declares 13 rows for mfcc but returns 39 (features.py:76-84,207-215)
    declared = sum(O.DECLARED_ROWS.get(k, 64) for k in idx)
    assert declared == int(g['feat_multi2_n_features'][0]) == 269
    assert out.shape[0] == 295


def test_feature_concat_quirk():
    """features.py:113 concatenates on dim 0 even for batched input."""
    g = golden()
    filters, _, _ = O.mel_filterbank()
    su, sb = crandn((2, 257, 30), 400).numpy(), crandn((4, 2, 257, 30), 401).numpy()
    names = ['logfbe', 'fbe', 'cubicpdf']
    out, idx = O.extract_features(su, filters, names)
    assert_parity(out, g['feat_multi_u'], 5e-6)
    assert [idx[k] for k in sorted(idx)] == [tuple(r) for r in g['feat_multi_u_idx']]
    out, idx = O.extract_features(sb, filters, names)
    assert out.shape == (12, 64, 30)
    assert_parity(out, g['feat_multi_b'], 5e-6)
    assert [idx[k] for k in sorted(idx)] == [tuple(r) for r in g['feat_multi_b_idx']]
    with pytest.raises(ValueError):
        O.extract_features(su, filters, ['nope'])
    with pytest.raises(ValueError):
        O.fbe(su[0], filters)


def _ffnn_transform(sources, stacks, decimation):
    filters, _, _ = O.mel_filterbank()
    spec = O.stft(sources)
    mix, fg = spec
    bg = mix - fg
    feats = O.fbe(mix, filters, compression='log')
    x = O.decimate(O.stack(feats, stacks), decimation)
    labels = O.decimate(O.irm(np.abs(fg), np.abs(bg), filters), decimation)
    return np.concatenate([x, labels])


def test_ffnn_glue():
    g = golden()
    sources = (0.05 * randn((2, 2, 4000), 500)).numpy()
    assert_parity(_ffnn_transform(sources, 5, 1), g['ffnn_transform'], 2e-5)
    assert_parity(_ffnn_transform(sources, 3, 2), g['ffnn_transform_s3d2'], 2e-5)
    feats = randn((3, 64, 17), 501).numpy()
    assert np.array_equal(O.stack(feats, 5), g['ffnn_stack_b'])
    assert np.array_equal(O.stack(feats[0], 5), g['ffnn_stack_u'])
    stacked = O.stack(feats, 5)
    assert_parity(O.static_normalize(stacked, g['ffnn_static_mean'],
                                     g['ffnn_static_std']),
                  g['ffnn_static'], 1e-6)
    # E[x^2] - E[x]^2 cancels in the float32 reference (ffnn.py:199-201): its
    # own rounding is ~1e-4 here, the float64 oracle is the tie-breaker
    assert_parity(O.cumulative_normalize(stacked), g['ffnn_cumulative'], 3e-4)
    # enhance tail
    filters, _, scaling = O.mel_filterbank()
    mix = (0.05 * randn((3, 2, 4000), 504)).numpy()
    X = O.stft(mix)
    mask = torch.sigmoid(randn((3, 64, X.shape[-1]), 505)).numpy()
    ext = O.mel_backward(filters, scaling, mask)
    assert_parity(ext, g['ffnn_enh_mask_ext'], 2e-6)
    y = O.istft(X.mean(1) * ext)[..., :4000]
    assert_parity(y, g['ffnn_enh_out'], 1e-5)


def _crit_inputs():
    B, S, L = 5, 3, 2000
    est, ref = randn((B, S, L), 600), randn((B, S, L), 601)
    est = ref.roll(1, 1) * 0.7 + 0.3 * est
    return est, ref, torch.tensor(golden()['crit_lengths'])


def test_criteria_values():
    g = golden()
    est, ref, lengths = _crit_inputs()
    e, r, n = est.numpy(), ref.numpy(), lengths.numpy()
    assert np.allclose(O.snr(e, r, n), g['crit_snr'], rtol=0, atol=2e-5)
    assert np.allclose(O.snr(e, r, n), g['crit_snr_f64'], rtol=0, atol=1e-9)
    loss, perm = O.sisnr(e, r, n)
    assert np.allclose(loss, g['crit_sisnr'], rtol=0, atol=2e-5)
    assert np.allclose(loss, g['crit_sisnr_f64'], rtol=0, atol=1e-9)
    assert (perm != np.arange(3)).any()  # PIT really permutes here
    assert np.allclose(O.snr(e[:, 0], r[:, 0], n), g['crit_snr_2d'], atol=2e-5)
    e4 = e.reshape(5, 3, 2, 1000)[..., :900]
    r4 = r.reshape(5, 3, 2, 1000)[..., :900]
    assert np.allclose(O.snr(e4, r4, np.minimum(n, 900)), g['crit_snr_4d'],
                       atol=2e-5)
    close = (ref + 1e-4 * randn((5, 3, 2000), 602)).numpy()
    # float32 reference loses digits here; the fp64 oracle is the tie-breaker
    assert np.allclose(O.snr(close, r, n), g['crit_snr_close'], atol=2e-2)
    assert np.allclose(O.sisnr(close, r, n)[0], g['crit_sisnr_close'], atol=2e-2)
    # torch port
    assert np.allclose(P.snr(est, ref, lengths).numpy(), g['crit_snr'], atol=1e-5)
    assert np.allclose(P.sisnr(est, ref, lengths).numpy(), g['crit_sisnr'], atol=1e-5)


def test_criteria_gradients():
    g = golden()
    est, ref, lengths = _crit_inputs()
    e, r, n = est.numpy(), ref.numpy(), lengths.numpy()
    assert_parity(O.snr_grad(e, r, n), g['crit_snr_grad'], 2e-5)
    assert int(g['crit_sisnr_ref_backward_ok']) == 0  # documented defect
    assert_parity(O.sisnr_grad(e, r, n), g['crit_sisnr_grad'], 2e-5)


def test_batched_equals_single():
    """tests/test_losses.py:13-57 property on the oracle (reduced sizes)."""
    rng = np.random.default_rng(0)
    B, S, Lmax = 6, 3, 3000
    lengths = rng.integers(1500, Lmax, B)
    x = np.zeros((B, S, Lmax))
    y = np.zeros((B, S, Lmax))
    for i, n in enumerate(lengths):
        y[i, :, :n] = rng.standard_normal((S, n))
    x = y + rng.standard_normal(y.shape)  # padding of x is NOT zero
    for fn in (O.snr, lambda a, b, c: O.sisnr(a, b, c)[0]):
        batched = fn(x, y, lengths)
        single = np.array([fn(x[i:i + 1, :, :n], y[i:i + 1, :, :n], [n])[0]
                           for i, n in enumerate(lengths)])
        assert np.allclose(batched, single, rtol=0, atol=1e-10)


def test_mse_and_multiresyu_port_against_reference():
    """oracle/torch_port.py restatements of criterion.py:104-226 against the reference's own
    outputs (values and autograd)."""
    g = golden()
    est, ref, lengths = _crit_inputs()
    weight = torch.tensor(g['crit_mse_weight'])
    assert np.allclose(P.mse(est, ref, lengths).numpy(), g['crit_mse'], atol=1e-6)
    assert np.allclose(P.mse(est, ref, lengths, weight).numpy(), g['crit_mse_w'], atol=1e-6)
    assert np.allclose(P.mse(est[:, 0], ref[:, 0], lengths).numpy(), g['crit_mse_2d'], atol=1e-6)
    cest = torch.complex(est[..., :1000], est[..., 1000:])
    cref = torch.complex(ref[..., :1000], ref[..., 1000:])
    assert np.allclose(P.mse(cest, cref, lengths.clamp(max=1000), weight).numpy(),
                       g['crit_mse_complex'], atol=1e-6)
    kws = {'def': {}, 'multi': dict(frame_lengths=[512, 256], hop_lengths=[128, 128],
                                    time_domain_weight=0.3, spectral_weight=0.7),
           'si': dict(scale_invariant=True)}
    for name, kw in kws.items():
        assert np.allclose(P.multiresyu(est, ref, lengths, **kw).numpy(), g[f'crit_mry_{name}'],
                           rtol=1e-5)
        assert np.allclose(P.multiresyu(est[:, 0], ref[:, 0], lengths, **kw).numpy(),
                           g[f'crit_mry_{name}_2d'], rtol=1e-5)
    e = est.clone().requires_grad_(True)
    (P.multiresyu(e, ref, lengths) * weight).sum().backward()
    assert_parity(e.grad.numpy(), g['crit_mry_def_grad'], 1e-5)


# --------------------------------------------------------------------------- #
# round 2: pad_mode / center variants, MANNER's loss                           #
# --------------------------------------------------------------------------- #
PAD_CASES = [('reflect_512_128', dict(frame_length=512, hop_length=128, pad_mode='reflect'), (3, 3001), 700),
             ('reflect_256_64', dict(frame_length=256, hop_length=64, pad_mode='reflect', normalized=False), (2, 2, 1500), 701),
             ('nocenter_512_128', dict(frame_length=512, hop_length=128, center=False), (3, 3001), 702),
             ('reflect_nocenter_400', dict(frame_length=400, hop_length=100, n_fft=512, center=False, pad_mode='reflect'), (2, 2777), 703)]


@pytest.mark.parametrize('tag,kw,shape,seed', PAD_CASES)
def test_oracle_stft_padding_modes(tag, kw, shape, seed):
    g = golden()
    x = randn(shape, seed)
    spec = O.stft(x.numpy(), **kw)
    assert spec.shape == g[f'stft_{tag}'].shape
    assert_parity(spec, g[f'stft_{tag}'], 2e-6, tag)
    w = crandn(tuple(spec.shape), seed + 50).numpy()
    grad = O.stft_grad(w, shape[-1], **kw)
    assert_parity(grad, g[f'stft_{tag}_grad'], 1e-5, tag + ' gradient')


def test_oracle_manner_loss():
    g = golden()
    mx, my = 0.1 * randn((3, 8000), 710), 0.1 * randn((3, 8000), 711)
    my = 0.7 * mx + 0.3 * my
    my[2, 5000:] = 0.0
    sc, mag = O.manner_stft_loss(mx.numpy(), my.numpy())
    assert np.allclose(sc, g['manner_def_sc'], rtol=2e-5) and np.allclose(mag, g['manner_def_mag'], rtol=2e-5)
    sc, mag = O.manner_stft_loss(mx.numpy(), my.numpy(), fft_sizes=[512, 256], hop_sizes=[128, 64],
                                 win_lengths=[512, 200], factor_sc=0.5, factor_mag=1.0)
    assert np.allclose(sc, g['manner_small_sc'], rtol=2e-5) and np.allclose(mag, g['manner_small_mag'], rtol=2e-5)
