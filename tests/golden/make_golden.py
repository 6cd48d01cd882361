"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  It imports ``brever.modules`` and
``brever.criterion`` from ``/root/reference`` (read-only, no stubs needed) and
``brever.models.ffnn`` with ``sys.modules`` stubs for the optional third-party
packages that are not installed here, runs them on CPU in float32 on seeded
inputs, and stores inputs' seeds + outputs as ``.npz``.

Nothing at test time reads ``/root/reference``; tests rebuild the inputs from
the recorded seeds and compare against the stored outputs.
"""
import itertools
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def _stub_optional_packages():
    class _Meta(type):
        def __getattr__(cls, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return 0

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return _Meta(name, (), {'__init__': lambda self, *a, **k: None})

    for name in ['batch_pystoi', 'pesq', 'pesq._pesq', 'pesq.cypesq',
                 'pystoi', 'soundfile', 'sofa', 'matplotlib',
                 'matplotlib.pyplot', 'torch_ema', 'h5py', 'wandb']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = _Anything(name)
                mod.__path__ = []  # let `import pkg.sub` resolve to stubs
                sys.modules[name] = mod


def randn(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


def crandn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    re = torch.randn(*shape, generator=g)
    im = torch.randn(*shape, generator=g)
    return torch.complex(re, im)


def main():
    sys.path.insert(0, REF)
    from brever.criterion import MultiResYuLoss, mse, sisnr, snr
    from brever.modules import STFT, FeatureExtractor, MelFilterbank

    torch.set_num_threads(1)
    out = {}

    # ---- STFT: the reference's own test input (tests/test_modules.py:318-326)
    combos = list(itertools.product([256, 128], [1.0, 0.5], [1.0, 0.15],
                                    [False, True], [False, True]))
    x = randn((4096,), 42)
    for hop, c, s, normalized, onesided in combos:
        kw = dict(frame_length=512, hop_length=hop, compression_factor=c,
                  scale_factor=s, normalized=normalized, onesided=onesided)
        st = STFT(**kw)
        spec = st(x)
        key = f'rt_h{hop}_c{c}_s{s}_n{int(normalized)}_o{int(onesided)}'
        # full spectra only for a subset (size); round trips for all
        if onesided and (c, s) in ((1.0, 1.0), (0.5, 0.15)):
            out[key + '_spec'] = spec.numpy()
        if not onesided and hop == 256 and c == 1.0 and s == 1.0:
            out[key + '_spec'] = spec.numpy()
        out[key + '_back'] = st.backward(spec.clone()).numpy()

    # ---- STFT: shape / padding cases incl. ragged lengths and n_fft > frame
    cases = [
        (100, 512, 256, None, 'hann'), (4000, 512, 256, None, 'hann'),
        (777, 512, 128, None, 'hann'), (1000, 256, 128, None, 'hann'),
        (3001, 510, 128, None, 'hann'), (1600, 400, 100, 512, 'hann'),
        (2048, 512, 256, None, 'hamming'), (900, 256, 64, None, None),
        (512, 512, 256, None, 'hann'), (513, 512, 256, None, 'hann'),
    ]
    for i, (S, L, H, nfft, win) in enumerate(cases):
        xs = randn((2, S), 100 + i)
        st = STFT(frame_length=L, hop_length=H, window=win, n_fft=nfft)
        spec = st(xs)
        key = f'shape{i}'
        out[key + '_meta'] = np.array([S, L, H, nfft or L], dtype=np.int64)
        out[key + '_spec'] = spec.numpy()
        out[key + '_strides'] = np.array(spec.stride(), dtype=np.int64)
        out[key + '_back'] = st.backward(spec.clone()).numpy()
    # leading dims + un-normalised + compression (SGMSE-style, sgmse.py:79-87)
    xs = randn((2, 3, 2500), 200)
    st = STFT(frame_length=510, hop_length=128, normalized=False,
              compression_factor=0.5, scale_factor=0.15)
    spec = st(xs)
    out['sgmse_spec'] = spec.numpy()
    out['sgmse_back'] = st.backward(spec.clone()).numpy()
    mag, phase = st(xs, return_type='mag_phase')
    out['sgmse_mag'], out['sgmse_phase'] = mag.numpy(), phase.numpy()
    # TF-GridNet-style 256/128 un-normalised (tfgridnet.py:69-74)
    xs = randn((3, 2, 3000), 201)
    st = STFT(frame_length=256, hop_length=128, normalized=False)
    spec = st(xs)
    out['gridnet_spec'] = spec.numpy()
    out['gridnet_back'] = st.backward(spec[:, :1].clone()).numpy()
    # iSTFT of a spectrogram that is NOT an STFT (masked), bin-major contiguous
    spec = crandn((2, 257, 20), 202)
    st = STFT(frame_length=512, hop_length=128)
    out['istft_random'] = st.backward(spec.clone()).numpy()

    # ---- ConvSTFT (stft.py:201-319): the reference's own test input and parametrisation
    #      (tests/test_modules.py:329-352) plus batched / ragged shapes
    from brever.modules import ConvSTFT
    x = randn((4096,), 42)
    for hop, c, s, normalized in itertools.product([256, 128], [1.0, 0.5], [1.0, 0.15],
                                                   [False, True]):
        cs = ConvSTFT(frame_length=512, hop_length=hop, compression_factor=c,
                      scale_factor=s, normalized=normalized)
        spec = cs(x)
        key = f'conv_h{hop}_c{c}_s{s}_n{int(normalized)}'
        if (c, s) in ((1.0, 1.0), (0.5, 0.15)):
            out[key + '_spec'] = spec.numpy()
        out[key + '_back'] = cs.backward(spec.clone()).numpy()
    for i, (S, L, H) in enumerate([(100, 512, 256), (3001, 512, 128), (777, 256, 64),
                                   (2000, 256, 256), (1500, 128, 32)]):
        xs = randn((2, 3, S), 150 + i)
        cs = ConvSTFT(frame_length=L, hop_length=H)
        spec = cs(xs)
        out[f'convshape{i}_meta'] = np.array([S, L, H], dtype=np.int64)
        out[f'convshape{i}_spec'] = spec.numpy()
        out[f'convshape{i}_back'] = cs.backward(spec.clone()).numpy()
    cs = ConvSTFT(frame_length=512, hop_length=128)
    out['conv_random_back'] = cs.backward(crandn((2, 257, 20), 160)).numpy()

    # ---- mel filterbank constants (bit-exact)
    for tag, kw in [('mel512', {}), ('mel256', dict(n_fft=256)),
                    ('mel40', dict(n_filters=40, n_fft=400, fs=8000, fmax=4000))]:
        fb = MelFilterbank(**kw)
        out[tag + '_filters'] = fb.filters.numpy()
        out[tag + '_fc'] = fb.fc.numpy()
        out[tag + '_scaling'] = fb.scaling.numpy()
        out[tag + '_inverse'] = fb.inverse_filters.numpy()
    fb = MelFilterbank()
    pw = randn((2, 257, 12), 300).abs()
    out['mel_fwd'] = fb(pw).numpy()
    out['mel_bwd'] = fb.backward(randn((2, 64, 12), 301)).numpy()

    # ---- features (tests/test_features.py input shape)
    spec_u = crandn((2, 257, 30), 400)
    spec_b = crandn((4, 2, 257, 30), 401)
    for name in ['fbe', 'logfbe', 'cubicfbe', 'pdf', 'logpdf', 'cubicpdf']:
        fe = FeatureExtractor(features=[name], mel_fb=fb)
        out[f'feat_u_{name}'] = fe(spec_u).numpy()
        out[f'feat_b_{name}'] = fe(spec_b).numpy()
    # binaural cues and DCT features (features.py:203-296, :199-219); the 0.1x copies keep the
    # recursive IC spectra below torchaudio.lfilter's default clamp to [-1, 1]
    for name in ['ild', 'ipd', 'ic', 'mfcc', 'cubicmfcc', 'pdfcc']:
        fe = FeatureExtractor(features=[name], mel_fb=fb)
        out[f'feat_u_{name}'] = fe(spec_u).numpy()
        out[f'feat_b_{name}'] = fe(spec_b).numpy()
        out[f'feat_bs_{name}'] = fe(0.1 * spec_b).numpy()
    fe = FeatureExtractor(features=['ic'], mel_fb=fb, hop_length=64)
    out['feat_bs_ic_hop64'] = fe(0.1 * spec_b).numpy()
    fe = FeatureExtractor(features=['mfcc', 'ild', 'logfbe', 'ic', 'ipd'], mel_fb=fb)
    out['feat_multi2_u'] = fe(spec_u).numpy()
    out['feat_multi2_u_idx'] = np.array(
        [fe.indices[k] for k in sorted(fe.indices)], dtype=np.int64)
    out['feat_multi2_n_features'] = np.array([fe.n_features], dtype=np.int64)
    fe = FeatureExtractor(features=['logfbe', 'fbe', 'cubicpdf'], mel_fb=fb)
    out['feat_multi_u'] = fe(spec_u).numpy()
    out['feat_multi_u_idx'] = np.array(
        [fe.indices[k] for k in sorted(fe.indices)], dtype=np.int64)
    out['feat_multi_b'] = fe(spec_b).numpy()
    out['feat_multi_b_idx'] = np.array(
        [fe.indices[k] for k in sorted(fe.indices)], dtype=np.int64)

    # ---- FFNN glue: transform / stack / normalisers / irm (ffnn.py:77-203)
    _stub_optional_packages()
    from brever.models.ffnn.ffnn import (FFNN, CumulativeNormalizer,
                                         StaticNormalizer)
    torch.manual_seed(0)
    model = FFNN()
    sources = 0.05 * randn((2, 2, 4000), 500)
    tr = model.transform(sources)
    out['ffnn_transform'] = tr.numpy()
    model3 = FFNN(stacks=3, decimation=2)
    out['ffnn_transform_s3d2'] = model3.transform(sources).numpy()
    feats = randn((3, 64, 17), 501)
    out['ffnn_stack_b'] = model.stack(feats).numpy()
    out['ffnn_stack_u'] = model.stack(feats[0]).numpy()
    norm = StaticNormalizer(384)
    mean, std = randn((384, 1), 502), randn((384, 1), 503).abs() + 0.5
    norm.set_statistics(mean, std)
    stacked = model.stack(feats)
    out['ffnn_static_mean'], out['ffnn_static_std'] = mean.numpy(), std.numpy()
    out['ffnn_static'] = norm(stacked).numpy()
    out['ffnn_cumulative'] = CumulativeNormalizer()(stacked).numpy()
    # enhance tail (ffnn.py:105-110) with a synthetic mask in place of the MLP
    mix = 0.05 * randn((3, 2, 4000), 504)
    X = model.stft(mix)
    mask = torch.sigmoid(randn((3, 64, X.shape[-1]), 505))
    ext = model.mel_fb.backward(mask)
    y = model.stft.backward(X.mean(1) * ext)[..., :4000]
    out['ffnn_enh_mask_ext'] = ext.numpy()
    out['ffnn_enh_out'] = y.numpy()

    # ---- criteria
    B, S, L = 5, 3, 2000
    lengths = torch.tensor([2000, 1500, 1999, 801, 1200])
    est, ref = randn((B, S, L), 600), randn((B, S, L), 601)
    est = ref.roll(1, 1) * 0.7 + 0.3 * est  # make PIT pick a non-identity perm
    out['crit_lengths'] = lengths.numpy()
    out['crit_snr'] = snr(est, ref, lengths).numpy()
    out['crit_sisnr'] = sisnr(est, ref, lengths).numpy()
    out['crit_snr_2d'] = snr(est[:, 0], ref[:, 0], lengths).numpy()
    out['crit_snr_4d'] = snr(est.view(B, S, 2, L // 2)[..., :900],
                             ref.view(B, S, 2, L // 2)[..., :900],
                             lengths.clamp(max=900)).numpy()
    e = est.clone().requires_grad_(True)
    snr(e, ref, lengths).sum().backward()
    out['crit_snr_grad'] = e.grad.numpy()
    # high-SNR case (cancellation check): estimate = target + tiny error
    close = ref + 1e-4 * randn((B, S, L), 602)
    out['crit_snr_close'] = snr(close, ref, lengths).numpy()
    out['crit_sisnr_close'] = sisnr(close, ref, lengths).numpy()
    # float64 reference values
    out['crit_snr_f64'] = snr(est.double(), ref.double(), lengths).numpy()
    out['crit_sisnr_f64'] = sisnr(est.double(), ref.double(), lengths).numpy()
    # sisnr gradient: the reference's own sisnr cannot be back-propagated
    # (in-place `max_snr /= S` on the amax output, criterion.py:69-70); verify
    # that, then take the gradient from the out-of-place restatement in
    # oracle/torch_port.py whose FORWARD is checked equal to the reference here.
    sys.path.insert(0, os.path.join(HERE, '..', '..'))
    from oracle import torch_port
    e = est.clone().requires_grad_(True)
    try:
        sisnr(e, ref, lengths).sum().backward()
        out['crit_sisnr_ref_backward_ok'] = np.array(1)
        out['crit_sisnr_grad'] = e.grad.numpy()
    except RuntimeError:
        out['crit_sisnr_ref_backward_ok'] = np.array(0)
        e = est.clone().requires_grad_(True)
        port = torch_port.sisnr(e, ref, lengths)
        assert torch.equal(port.detach(), sisnr(est, ref, lengths))
        port.sum().backward()
        out['crit_sisnr_grad'] = e.grad.numpy()

    # ---- mse (criterion.py:104-132) and multiresyu (criterion.py:135-226)
    weight = torch.tensor([1.0, 0.5, 2.0, 0.25, 1.5])
    out['crit_mse_weight'] = weight.numpy()
    out['crit_mse'] = mse(est, ref, lengths).numpy()
    out['crit_mse_w'] = mse(est, ref, lengths, weight=weight).numpy()
    out['crit_mse_2d'] = mse(est[:, 0], ref[:, 0], lengths).numpy()
    cest = torch.complex(est[..., :1000], est[..., 1000:])
    cref = torch.complex(ref[..., :1000], ref[..., 1000:])
    clen = lengths.clamp(max=1000)
    out['crit_mse_complex'] = mse(cest, cref, clen, weight=weight).numpy()
    e = est.clone().requires_grad_(True)
    mse(e, ref, lengths, weight=weight).sum().backward()
    out['crit_mse_grad'] = e.grad.numpy()
    ce = cest.clone().requires_grad_(True)
    mse(ce, cref, clen).sum().backward()
    out['crit_mse_complex_grad'] = ce.grad.numpy()
    for name, kw in (('def', {}),
                     ('multi', dict(frame_lengths=[512, 256], hop_lengths=[128, 128],
                                    time_domain_weight=0.3, spectral_weight=0.7)),
                     ('si', dict(scale_invariant=True))):
        crit = MultiResYuLoss(**kw)
        out[f'crit_mry_{name}'] = crit(est, ref, lengths).numpy()
        out[f'crit_mry_{name}_2d'] = crit(est[:, 0], ref[:, 0], lengths).numpy()
        if name != 'si':
            e = est.clone().requires_grad_(True)
            (crit(e, ref, lengths) * weight).sum().backward()
            out[f'crit_mry_{name}_grad'] = e.grad.numpy()

    # ---- round 2: pad_mode / center variants of STFT.forward (stft.py:59-77,140-144) ----------
    for tag, kw, shape, seed in [('reflect_512_128', dict(frame_length=512, hop_length=128, pad_mode='reflect'), (3, 3001), 700),
                                 ('reflect_256_64', dict(frame_length=256, hop_length=64, pad_mode='reflect', normalized=False), (2, 2, 1500), 701),
                                 ('nocenter_512_128', dict(frame_length=512, hop_length=128, center=False), (3, 3001), 702),
                                 ('reflect_nocenter_400', dict(frame_length=400, hop_length=100, n_fft=512, center=False, pad_mode='reflect'), (2, 2777), 703)]:
        st = STFT(**kw)
        x = randn(shape, seed)
        spec = st(x)
        out[f'stft_{tag}'] = spec.numpy()
        w = crandn(tuple(spec.shape), seed + 50)
        xg = x.clone().requires_grad_(True)
        sg = st(xg)
        (sg.real * w.real + sg.imag * w.imag).sum().backward()
        out[f'stft_{tag}_grad'] = xg.grad.numpy()
        if kw.get('center', True):
            out[f'stft_{tag}_back'] = st.backward(spec.clone()).numpy()

    # ---- round 2: MANNER multi-resolution STFT loss (models/manner/stft_loss.py:22-151) --------
    from brever.models.manner.stft_loss import MultiResolutionSTFTLoss
    mx, my = 0.1 * randn((3, 8000), 710), 0.1 * randn((3, 8000), 711)
    my = 0.7 * mx + 0.3 * my
    my[2, 5000:] = 0.0                                       # silence: the clamp is active there
    for tag, kw in (('def', {}), ('small', dict(fft_sizes=[512, 256], hop_sizes=[128, 64], win_lengths=[512, 200],
                                                factor_sc=0.5, factor_mag=1.0))):
        crit = MultiResolutionSTFTLoss(**kw)
        e = mx.clone().requires_grad_(True)
        sc, mag = crit(e, my)
        out[f'manner_{tag}_sc'], out[f'manner_{tag}_mag'] = sc.detach().numpy(), mag.detach().numpy()
        wsc, wmag = torch.tensor([1.0, -0.5, 2.0]), torch.tensor([0.3, 1.5, -1.0])
        ((sc * wsc).sum() + (mag * wmag).sum()).backward()
        out[f'manner_{tag}_grad'] = e.grad.numpy()

    # ---- round 2: per-utterance transform + collate (training.py:336-338, data.py:408-491) ----
    batch = 0.05 * randn((3, 2, 2, 4000), 720)
    blen = [4000, 3000, 2345]
    for i, n in enumerate(blen):
        batch[i, ..., n:] = 0
    for tag, mdl in (('def', model), ('s3d2', model3)):
        items = [mdl.transform(batch[i, ..., :n]) for i, n in enumerate(blen)]
        tmax = max(t.shape[-1] for t in items)
        out[f'ffnn_tbatch_{tag}'] = torch.stack(
            [torch.nn.functional.pad(t, (0, tmax - t.shape[-1])) for t in items]).numpy()
        out[f'ffnn_tbatch_{tag}_len'] = np.array([t.shape[-1] for t in items], dtype=np.int64)
    sg_stft = STFT(frame_length=512, hop_length=128, window='hann', compression_factor=0.5,
                   scale_factor=0.15, normalized=False)
    items = []
    for i, n in enumerate(blen):
        src = batch[i, ..., :n].clone().mean(axis=-2)        # sgmse.py:153-158
        src /= src[0].abs().max()
        items.append(sg_stft(src)[..., :-1, :])
    tmax = max(t.shape[-1] for t in items)
    out['sgmse_tbatch'] = torch.stack(
        [torch.nn.functional.pad(t, (0, tmax - t.shape[-1])) for t in items]).numpy()
    out['sgmse_tbatch_len'] = np.array([t.shape[-1] for t in items], dtype=np.int64)

    # ---- round 2: gradient of the scale-invariant MultiResYuLoss (criterion.py:207-226) --------
    crit = MultiResYuLoss(scale_invariant=True)
    e = est.clone().requires_grad_(True)
    (crit(e, ref, lengths) * weight).sum().backward()
    out['crit_mry_si_grad'] = e.grad.numpy()

    path = os.path.join(HERE, 'reference_vectors.npz')
    if os.path.exists(path):      # regenerating must not move any vector already committed
        prev = np.load(path)
        moved = [k for k in prev.files if k not in out or not np.array_equal(prev[k], out[k], equal_nan=True)]
        print('arrays that differ from the committed file:', moved or 'none')
    np.savez_compressed(path, **out)
    total = sum(v.nbytes for v in out.values())
    print(f'wrote {len(out)} arrays, {total / 1e6:.2f} MB raw')


if __name__ == '__main__':
    main()
