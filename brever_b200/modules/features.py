"""Drop-in ``FeatureExtractor`` (brever/modules/features.py:13-220).

The filterbank-energy family (``fbe``, ``logfbe``, ``cubicfbe``, ``pdf``,
``logpdf``, ``cubicpdf``) — ``logfbe`` is the FFNN default (ffnn.py:20) — runs
in one fused kernel (``brv_fbe_features``): |X|^2, channel mean, banded mel,
optional pdf normalisation and compression, never materialising the magnitude
or power spectrograms the reference creates (features.py:186-190).

The binaural cues (``ild``, ``ipd``, ``ic``; features.py:222-296) and the DCT
features (``mfcc``, ``cubicmfcc``, ``pdfcc``; features.py:199-219) run through the
same kernel with another per-bin quantity in its first phase / a DCT stage after
the compression (``brv_mel_features``); ``ic`` first runs the recursive auto /
cross spectra along frames (``brv_ic_coherence``).  There is no PyTorch or CPU
fallback behind this class.
"""
import math

import torch

from .. import _lib

eps = torch.finfo().eps  # features.py:10 (float32 eps, 1.1920929e-07)

_COMPRESSION = {'none': 0, 'log': 1, 'cubic': 2}
_SQRT = 3                       # kernel-side compression code used by ic
_MODE_POWER, _MODE_ILD, _MODE_IPD, _MODE_REAL = 0, 1, 2, 3


class FeatureExtractor:
    # name -> (normalize, compression); features.py:21-101
    _FBE_FAMILY = {
        'fbe': (False, 'none'),
        'logfbe': (False, 'log'),
        'cubicfbe': (False, 'cubic'),
        'pdf': (True, 'none'),
        'logpdf': (True, 'log'),
        'cubicpdf': (True, 'cubic'),
    }
    # name -> (normalize, compression) with dct=True; features.py:76-100
    _DCT_FAMILY = {
        'mfcc': (False, 'log'),
        'cubicmfcc': (False, 'cubic'),
        'pdfcc': (True, 'log'),
    }

    def __init__(self, features, mel_fb, hop_length=256, fs=16e3):
        self.features = sorted(features)
        self.mel_fb = mel_fb
        self.hop_length = hop_length
        self.fs = fs
        self.indices = None
        self._dct = {}

    def _feature_count(self, feature):
        # what the reference *declares* (features.py:21-101): the DCT features say 13 but
        # return 39 rows (deltas and double deltas default to True) -- kept as is
        if feature in self._FBE_FAMILY or feature in ('ild', 'ipd', 'ic'):
            return self.mel_fb.n_filters
        if feature in self._DCT_FAMILY:
            return 13
        raise ValueError(f'unrecognized feature, got {feature}')

    @property
    def n_features(self):
        return sum(self._feature_count(f) for f in self.features)

    def __call__(self, x):
        # features.py:103-113 — sorted names, concatenated along dim 0 (which
        # is the BATCH dim for batched input; quirk kept for parity)
        output = []
        self.indices = {}
        i_start = 0
        for feature in self.features:
            data = self.calc_feature(x, feature)
            output.append(data)
            i_end = i_start + len(data)
            self.indices[feature] = (i_start, i_end)
            i_start = i_end
        return torch.cat(output)

    def calc_feature(self, x, feature):
        unbatched = x.ndim == 3
        if unbatched:
            x = x.unsqueeze(0)
        elif x.ndim != 4:
            raise ValueError(f'input must be 3 or 4 dimensional, got {x.ndim}')
        if feature in self._FBE_FAMILY:
            normalize, compression = self._FBE_FAMILY[feature]
            out = self.fbe(x, normalize=normalize, compression=compression)
        elif feature in self._DCT_FAMILY:
            normalize, compression = self._DCT_FAMILY[feature]
            out = self.fbe(x, normalize=normalize, compression=compression, dct=True)
        elif feature == 'ild':
            out = self.ild(x)
        elif feature == 'ipd':
            out = self.ipd(x)
        elif feature == 'ic':
            out = self.ic(x)
        else:
            raise ValueError(f'unrecognized feature, got {feature}')
        return out.squeeze(0) if unbatched else out

    def fbe(self, x, normalize=False, compression='none', dct=False, n_dct=14,
            dct_type=2, dct_norm='ortho', return_dc=False, return_deltas=True,
            return_double_deltas=True, stacks=0, decimation=1, mean=None, std=None):
        """Filterbank energies of a ``(B, C, F, T)`` complex STFT -> ``(B, M, T)``.

        ``stacks`` / ``decimation`` / ``mean`` / ``std`` additionally fuse
        ``FFNN.stack``, ``FFNN.decimate`` and ``StaticNormalizer`` into the same
        pass (ffnn.py:122-135,186-187) -> ``(B, M*(stacks+1), ceil(T/dec))``.
        """
        if compression not in _COMPRESSION:
            raise ValueError('compression must be log, cubic or none, got '
                             f'{compression}')
        if dct:
            if (dct_type, dct_norm, return_dc, bool(return_deltas),
                    bool(return_double_deltas)) != (2, 'ortho', False, True, True):
                raise NotImplementedError(
                    'only the DCT configuration the reference features use is built: type 2, '
                    "norm 'ortho', no DC term, deltas and double deltas (features.py:76-100)")
            if stacks or decimation != 1 or mean is not None or std is not None:
                raise ValueError('the DCT features are not fused with stacking')
            return self._mel_features(x, _MODE_POWER, normalize, _COMPRESSION[compression],
                                      n_dct=n_dct - 1)
        _lib.require_cuda(x, 'FeatureExtractor input')
        if not x.is_complex():
            raise RuntimeError('FeatureExtractor input must be a complex STFT')
        if x.ndim != 4:
            raise ValueError(f'input must be 4 dimensional, got {x.ndim}')
        if x.dtype != torch.complex64:
            x = x.to(torch.complex64)
        x = x.resolve_conj().resolve_neg()   # raw pointers below: no lazy conj / neg bits
        batch, channels, bins, frames = x.shape
        fb = self.mel_fb
        vals, cols, rowptr, n_mel, n_in = fb.csr('forward', x.device)
        if bins != n_in:
            raise RuntimeError(f'expected {n_in} frequency bins, got {bins}')
        rows = n_mel * (stacks + 1)
        out_frames = -(-frames // decimation)
        out = torch.empty((batch, rows, out_frames), dtype=torch.float32,
                          device=x.device)

        def stat(t):
            if t is None:
                return None
            t = t.detach().to(device=x.device, dtype=torch.float32).reshape(-1)
            if t.numel() != rows:
                raise RuntimeError(f'statistics must have {rows} entries')
            return t.contiguous()
        mean_t, std_t = stat(mean), stat(std)
        with _lib.on_device(x.device):
            _lib.check(_lib.lib().brv_fbe_features(
                _lib.ptr(x), x.stride(0), x.stride(1), x.stride(2), x.stride(3),
                batch, channels, bins, frames, _lib.ptr(vals), _lib.ptr(cols),
                _lib.ptr(rowptr), n_mel, int(vals.numel()), int(normalize),
                _COMPRESSION[compression], float(eps), int(stacks),
                int(decimation), _lib.ptr(mean_t), _lib.ptr(std_t),
                _lib.ptr(out), _lib.stream_ptr(x.device)))
        return out

    # -- binaural cues and the DCT stage (features.py:199-296) ----------------------
    def _dct_basis(self, n_dct, device):
        """Rows 1..n_dct of the orthonormal DCT-II matrix over the mel bands
        (scipy.fft.dct(type=2, norm='ortho'), features.py:203-206)."""
        key = (n_dct, str(device))
        if key not in self._dct:
            m = self.mel_fb.n_filters
            k = torch.arange(1, n_dct + 1, dtype=torch.float64)[:, None]
            n = torch.arange(m, dtype=torch.float64)[None, :]
            basis = math.sqrt(2.0 / m) * torch.cos(math.pi * k * (2 * n + 1) / (2 * m))
            self._dct[key] = basis.to(torch.float32).contiguous().to(device)
        return self._dct[key]

    def __getstate__(self):          # device constants are rebuilt after unpickling
        state = dict(self.__dict__)
        state['_dct'] = {}
        return state

    def _mel_features(self, x, mode, normalize, compression, n_dct=0, real=None):
        """One launch of ``brv_mel_features``: ``(B, C, F, T)`` complex (or, for mode 3,
        a real ``(B, T, F)`` map) -> ``(B, n_mel | 3*n_dct, T)`` float32."""
        fb = self.mel_fb
        if real is None:
            _lib.require_cuda(x, 'FeatureExtractor input')
            if not x.is_complex():
                raise RuntimeError('FeatureExtractor input must be a complex STFT')
            if x.ndim != 4:
                raise ValueError(f'input must be 4 dimensional, got {x.ndim}')
            if x.dtype != torch.complex64:
                x = x.to(torch.complex64)
            x = x.resolve_conj().resolve_neg()   # raw pointers below: no lazy conj / neg bits
            batch, channels, bins, frames = x.shape
            src, strides = x, (x.stride(0), x.stride(1), x.stride(2), x.stride(3))
        else:
            batch, frames, bins = real.shape
            channels = 1
            src, strides = real, (real.stride(0), 0, real.stride(2), real.stride(1))
        device = src.device
        vals, cols, rowptr, n_mel, n_in = fb.csr('forward', device)
        if bins != n_in:
            raise RuntimeError(f'expected {n_in} frequency bins, got {bins}')
        if mode in (_MODE_ILD, _MODE_IPD) and channels < 2:
            raise IndexError('index 1 is out of bounds for dimension 1 with size 1')
        basis = self._dct_basis(n_dct, device) if n_dct else None
        rows = 3 * n_dct if n_dct else n_mel
        out = torch.empty((batch, rows, frames), dtype=torch.float32, device=device)
        with _lib.on_device(device):
            _lib.check(_lib.lib().brv_mel_features(
                _lib.ptr(src), *strides, batch, channels, bins, frames, mode,
                _lib.ptr(vals), _lib.ptr(cols), _lib.ptr(rowptr), n_mel,
                int(vals.numel()), int(normalize), int(compression), float(eps),
                _lib.ptr(basis), int(n_dct), _lib.ptr(out), _lib.stream_ptr(device)))
        return out

    def ild(self, x):
        """Interaural level difference, features.py:222-240: ``(B, 2, F, T)`` -> ``(B, M, T)``."""
        return self._mel_features(x, _MODE_ILD, False, 0)

    def ipd(self, x):
        """Interaural phase difference, features.py:242-260 (not wrapped to (-pi, pi])."""
        return self._mel_features(x, _MODE_IPD, False, 0)

    def ic(self, x, tau=10e-3):
        """Interaural coherence, features.py:262-296."""
        _lib.require_cuda(x, 'FeatureExtractor input')
        if not x.is_complex():
            raise RuntimeError('FeatureExtractor input must be a complex STFT')
        if x.ndim != 4:
            raise ValueError(f'input must be 4 dimensional, got {x.ndim}')
        if x.dtype != torch.complex64:
            x = x.to(torch.complex64)
        x = x.resolve_conj().resolve_neg()   # raw pointers below: no lazy conj / neg bits
        batch, channels, bins, frames = x.shape
        if channels < 2:
            raise IndexError('index 1 is out of bounds for dimension 1 with size 1')
        alpha = math.exp(-self.hop_length / (tau * self.fs))
        # the float32 coefficients the reference builds: tensor([1, -alpha]), tensor([1 - alpha, 0])
        a1 = torch.tensor(-alpha, dtype=torch.float32).item()
        b0 = torch.tensor(1 - alpha, dtype=torch.float32).item()
        coh = torch.empty((batch, frames, bins), dtype=torch.float32, device=x.device)
        with _lib.on_device(x.device):
            _lib.check(_lib.lib().brv_ic_coherence(
                _lib.ptr(x), x.stride(0), x.stride(1), x.stride(2), x.stride(3), batch,
                channels, bins, frames, b0, a1, _lib.ptr(coh), _lib.stream_ptr(x.device)))
        return self._mel_features(None, _MODE_REAL, False, _SQRT, real=coh)
