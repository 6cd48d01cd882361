"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/brever_b200.h declares, the host-side mirror keeps the
reference's names / defaults / error behaviour, init-time constants are
bit-identical to the reference's, and nothing silently falls back to CPU."""
import inspect
import os
import pickle
import re

import numpy as np
import pytest
import torch

import brever_b200 as brv
from brever_b200 import _lib
from oracle import tf_oracle as O

from _util import STFT_SHAPE_CASES, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'brever_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(brv_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 20
    lib = _lib.lib()
    for name in names:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(names) == set(_lib.PROTOTYPES), \
        set(names).symmetric_difference(_lib.PROTOTYPES)
    assert lib.brv_abi_version() == 1
    assert lib.brv_status_string(-4).decode() == 'window overlap add min: 1'


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, 'brever_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src, os.path.join(dirpath, f)


def test_cpu_tensors_fail_loudly():
    stft = brv.STFT()
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        stft(torch.randn(4096))
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        stft.backward(torch.randn(257, 9, dtype=torch.complex64))
    fb = brv.MelFilterbank()
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        fb(torch.rand(257, 4))
    fe = brv.FeatureExtractor({'logfbe'}, fb)
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        fe(torch.randn(2, 257, 4, dtype=torch.complex64))
    x = torch.randn(2, 1, 100)
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        brv.snr(x, x, torch.tensor([100, 100]))
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        brv.sisnr(x, x, torch.tensor([100, 100]))


def test_stft_signature_and_attributes():
    sig = inspect.signature(brv.STFT.__init__)
    assert list(sig.parameters)[1:] == [
        'frame_length', 'hop_length', 'window', 'center', 'pad_mode',
        'normalized', 'onesided', 'compression_factor', 'scale_factor', 'n_fft']
    defaults = {k: v.default for k, v in sig.parameters.items() if k != 'self'}
    assert defaults == dict(frame_length=512, hop_length=256, window='hann',
                            center=True, pad_mode='constant', normalized=True,
                            onesided=True, compression_factor=1, scale_factor=1,
                            n_fft=None)
    st = brv.STFT()
    assert (st.frame_length, st.hop_length, st.n_fft) == (512, 256, 512)
    assert st.window.dtype == torch.float64 and st.window.shape == (512,)
    assert abs(float(st.window.pow(2).sum()) - 192.0) < 1e-12
    assert np.array_equal(st.window.numpy(), O.get_window('hann', 512)) or \
        np.allclose(st.window.numpy(), O.get_window('hann', 512), atol=1e-16)
    assert brv.STFT(window=None).window.eq(1).all()
    assert brv.STFT(400, 100, n_fft=512).n_fft == 512
    with pytest.raises(ValueError):
        st(torch.zeros(10), return_type='nope')
    with pytest.raises(ValueError):
        st.backward(torch.zeros(10), input_type='nope')
    # plain Python object, picklable, no nn.Module state
    assert not isinstance(st, torch.nn.Module)
    clone = pickle.loads(pickle.dumps(st))
    assert clone.frame_length == 512 and clone._plans == {}


@pytest.mark.parametrize('case', STFT_SHAPE_CASES)
def test_frame_arithmetic_matches_reference(case):
    """Integer frame arithmetic: bit-exact against shapes the reference produced."""
    S, L, H, nfft, win = case
    i = STFT_SHAPE_CASES.index(case)
    ref = golden()[f'shape{i}_spec']
    st = brv.STFT(frame_length=L, hop_length=H, window=win, n_fft=nfft)
    assert st.n_bins == ref.shape[1]
    assert st.n_frames(S) == ref.shape[2] == O.stft_frames(S, L, H, nfft)
    assert st.frame_count(S) == O.frame_count(S, L, H)
    assert st.pad(torch.zeros(S)).shape[-1] == S + O.right_padding(S, L, H)


def test_frame_arithmetic_known_values():
    st = brv.STFT(512, 256)
    assert st.n_frames(64000) == 251 and st.n_frames(63999) == 251
    assert st.n_frames(100) == 3 and st.n_frames(4000) == 17
    assert brv.STFT(512, 128).n_frames(64000) == 501
    assert brv.STFT(510, 128).n_frames(128000) == 1001
    assert brv.STFT(510, 128).n_bins == 256
    assert brv.STFT(256, 128).n_frames(64000) == 501
    assert brv.STFT(400, 100, n_fft=512).n_frames(16000) == 161


@pytest.mark.parametrize('tag,kw', [
    ('mel512', {}), ('mel256', dict(n_fft=256)),
    ('mel40', dict(n_filters=40, n_fft=400, fs=8000, fmax=4000))])
def test_mel_constants_bit_identical_to_reference(tag, kw):
    g = golden()
    fb = brv.MelFilterbank(**kw)
    assert np.array_equal(fb.filters.numpy(), g[tag + '_filters'])
    assert np.array_equal(fb.fc.numpy(), g[tag + '_fc'])
    assert np.array_equal(fb.scaling.numpy(), g[tag + '_scaling'])
    assert np.array_equal(fb.inverse_filters.numpy(), g[tag + '_inverse'])
    assert fb.n_filters == kw.get('n_filters', 64)
    pickle.loads(pickle.dumps(fb))


def test_feature_extractor_surface():
    fb = brv.MelFilterbank()
    fe = brv.FeatureExtractor({'logfbe', 'ild', 'mfcc'}, fb)
    assert fe.features == ['ild', 'logfbe', 'mfcc']
    assert fe.n_features == 64 + 64 + 13
    assert fe.indices is None
    with pytest.raises(ValueError, match='unrecognized feature'):
        brv.FeatureExtractor({'nope'}, fb).n_features
    with pytest.raises(ValueError, match='3 or 4 dimensional'):
        fe.calc_feature(torch.zeros(257, 4, dtype=torch.complex64), 'logfbe')


def test_conv_stft_host_surface():
    """ConvSTFT (stft.py:201-319): constructor state, padding arithmetic and the kernel bank are
    host-side and must match the reference / the oracle without a GPU."""
    g = golden()
    cs = brv.ConvSTFT(frame_length=512, hop_length=128)
    assert cs.frame_length == 512 and cs.hop_length == 128 and cs.normalized
    assert cs._normalization_factor == 0.5 * 512 / 128 ** 0.5
    assert torch.allclose(cs.window.pow(2), torch.from_numpy(O.get_window('hann', 512)))
    bank, factor = O._conv_filters(512, 128, 'hann', True)
    ref = np.concatenate([bank.real, bank.imag])[:, None, :]
    assert cs.filters.shape == (514, 1, 512)
    assert np.abs(cs.filters.numpy() - ref).max() < 1e-7 and factor == cs._normalization_factor
    for i in range(5):
        S, L, H = (int(v) for v in g[f'convshape{i}_meta'])
        c = brv.ConvSTFT(frame_length=L, hop_length=H)
        assert c.n_frames(S) == g[f'convshape{i}_spec'].shape[-1]
        assert c.frame_count(S) == O.frame_count(S, L, H)
        assert c.pad(torch.zeros(S)).shape[-1] == (c.n_frames(S) - 1) * H + L + ((S + O.right_padding(S, L, H) + 2 * (L - H) - L) % H)
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        cs(torch.zeros(4096))
    pickle.loads(pickle.dumps(cs))


def test_registry_contract():
    assert set(brv.CriterionRegistry.keys()) >= {'sisnr', 'snr'}
    assert brv.init_criterion('snr') is brv.snr
    assert brv.init_criterion('sisnr') is brv.sisnr
    with pytest.raises(KeyError):
        brv.CriterionRegistry.get('nope')
    with pytest.raises(ValueError):
        brv.CriterionRegistry.register('snr')(lambda: None)
    reg = brv.Registry('thing')

    @reg.register('cls')
    class Thing:
        def __init__(self, a=1):
            self.a = a
    assert reg.get('cls') is Thing


def test_criterion_shape_asserts():
    x = torch.zeros(2, 3, 10)
    with pytest.raises(AssertionError):
        brv.sisnr(x, torch.zeros(2, 3, 11), torch.tensor([10, 10]))
    with pytest.raises(AssertionError):
        brv.sisnr(x[0], x[0], torch.tensor([10, 10]))
    with pytest.raises(AssertionError):
        brv.snr(x[0, 0], x[0, 0], torch.tensor([10]))
    with pytest.raises(AssertionError):
        brv.apply_mask(x, x, torch.tensor([10]))


def test_static_normalizer_state_dict_matches_reference_layout():
    norm = brv.ffnn.StaticNormalizer(384)
    sd = norm.state_dict()
    assert set(sd) == {'mean', 'std'}
    assert sd['mean'].shape == (384, 1) and sd['std'].shape == (384, 1)
    front = brv.ffnn.FFNNFrontEnd()
    assert front.input_size == 384


def test_metric_wrappers_mirror_reference_checks():
    """brever/metrics.py:112-150: registry names, `_check_input` shapes / defaults / errors.  When
    the reference tree is present its own `_check_input` source is executed next to ours."""
    import ast
    import os
    import torch
    from brever_b200 import metrics as M
    assert set(M.MetricRegistry.keys()) == {'snr', 'sisnr'}
    x, y = torch.zeros(3, 50), torch.zeros(3, 50)
    fx, fy, lengths, unbatched = M._check_input(x, y, None)
    assert fx.shape == (3, 1, 50) and fy.shape == (3, 1, 50) and not unbatched
    assert lengths.tolist() == [50, 50, 50]
    fx, _, lengths, unbatched = M._check_input(x[0], y[0], None)
    assert fx.shape == (1, 1, 50) and unbatched and lengths.tolist() == [50]
    cases = [(x, y[:2], None), (x[None], y[None], None), (x, y, [50, 50]), (x, y, [50, 51, 2])]
    for a, b, ln in cases:
        with pytest.raises(ValueError):
            M._check_input(a, b, ln)
    ref_path = '/root/reference/brever/metrics.py'
    if os.path.exists(ref_path):
        tree = ast.parse(open(ref_path).read())
        fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == '_check_input'][0]
        ns = {'torch': torch}
        exec(compile(ast.Module([fn], []), ref_path, 'exec'), ns)
        for a, b, ln in cases:
            with pytest.raises(ValueError) as e_ref:
                ns['_check_input'](a, b, ln)
            with pytest.raises(ValueError) as e_new:
                M._check_input(a, b, ln)
            assert str(e_ref.value) == str(e_new.value)
        for a, b, ln in [(x, y, None), (x[0], y[0], None), (x, y, torch.tensor([50, 10, 3]))]:
            r, n = ns['_check_input'](a, b, ln), M._check_input(a, b, ln)
            assert r[0].shape == n[0].shape and r[3] == n[3] and list(r[2]) == list(n[2])
