"""Drop-in ``FeatureExtractor`` (brever/modules/features.py:13-220).

The filterbank-energy family (``fbe``, ``logfbe``, ``cubicfbe``, ``pdf``,
``logpdf``, ``cubicpdf``) — ``logfbe`` is the FFNN default (ffnn.py:20) — runs
in one fused kernel (``brv_fbe_features``): |X|^2, channel mean, banded mel,
optional pdf normalisation and compression, never materialising the magnitude
or power spectrograms the reference creates (features.py:186-190).

The binaural cues (``ild``, ``ipd``, ``ic``) and the DCT features (``mfcc``,
``cubicmfcc``, ``pdfcc``) are SURVEY.md §8(f) rank-3 "next" rows and raise
``NotImplementedError`` until their kernels exist — there is no PyTorch or CPU
fallback behind this class.
"""
import torch

from .. import _lib

eps = torch.finfo().eps  # features.py:10 (float32 eps, 1.1920929e-07)

_COMPRESSION = {'none': 0, 'log': 1, 'cubic': 2}


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
    _NOT_BUILT = {'ild': None, 'ipd': None, 'ic': None, 'mfcc': 13,
                  'cubicmfcc': 13, 'pdfcc': 13}

    def __init__(self, features, mel_fb, hop_length=256, fs=16e3):
        self.features = sorted(features)
        self.mel_fb = mel_fb
        self.hop_length = hop_length
        self.fs = fs
        self.indices = None

    def _feature_count(self, feature):
        if feature in self._FBE_FAMILY:
            return self.mel_fb.n_filters
        if feature in self._NOT_BUILT:
            n = self._NOT_BUILT[feature]
            return self.mel_fb.n_filters if n is None else n
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
        elif feature in self._NOT_BUILT:
            raise NotImplementedError(
                f'feature "{feature}" has no sm_100a kernel yet (SURVEY.md '
                '§8f); brever_b200 has no PyTorch fallback')
        else:
            raise ValueError(f'unrecognized feature, got {feature}')
        return out.squeeze(0) if unbatched else out

    def fbe(self, x, normalize=False, compression='none', dct=False,
            stacks=0, decimation=1, mean=None, std=None):
        """Filterbank energies of a ``(B, C, F, T)`` complex STFT -> ``(B, M, T)``.

        ``stacks`` / ``decimation`` / ``mean`` / ``std`` additionally fuse
        ``FFNN.stack``, ``FFNN.decimate`` and ``StaticNormalizer`` into the same
        pass (ffnn.py:122-135,186-187) -> ``(B, M*(stacks+1), ceil(T/dec))``.
        """
        if compression not in _COMPRESSION:
            raise ValueError('compression must be log, cubic or none, got '
                             f'{compression}')
        if dct:
            raise NotImplementedError('DCT features have no sm_100a kernel yet')
        _lib.require_cuda(x, 'FeatureExtractor input')
        if not x.is_complex():
            raise RuntimeError('FeatureExtractor input must be a complex STFT')
        if x.ndim != 4:
            raise ValueError(f'input must be 4 dimensional, got {x.ndim}')
        if x.dtype != torch.complex64:
            x = x.to(torch.complex64)
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
