"""CPU oracle for the time-frequency front-end hot path (TEST INFRASTRUCTURE ONLY).

This module is a numpy restatement of the algorithms behind the reference's
``brever.modules.STFT``, ``MelFilterbank``, ``FeatureExtractor.fbe``, the FFNN
feature glue and the ``snr`` / ``sisnr`` criteria.  It exists to CHECK the CUDA
path; nothing under ``brever_b200/`` imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.

Where the arithmetic lives
--------------------------
The reference delegates the transforms to un-vendored third-party code:
``torch.stft`` / ``torch.istft`` (``torch==2.1.1``, requirements.txt:19) called at
``brever/modules/stft.py:66-77,126-136`` and ``scipy.signal.get_window``
(``scipy==1.11.4``, requirements.txt:17) at ``stft.py:49``.  Their published
algorithms are restated here explicitly (framing, periodic windows, real DFT,
overlap-add with window-sum-square normalisation) in float64.

Pinning
-------
PINNED: ``tests/golden/*.npz`` hold outputs of the reference itself, imported
from ``/root/reference`` in the build container by ``tests/golden/make_golden.py``
(committed).  ``tests/test_oracle_golden.py`` checks every function below against
those vectors, and against the reference's own round-trip / batched-vs-single
properties (``tests/test_modules.py:300-326``, ``tests/test_losses.py:13-57``).

All functions take/return numpy arrays; complex spectra are ``(..., F, T)``
like the reference (bins, then frames).
"""
import itertools
import math

import numpy as np

EPS32 = float(np.finfo(np.float32).eps)  # criterion.py:9, features.py:10
EPS64 = float(np.finfo(np.float64).eps)  # ffnn.py:12


# --------------------------------------------------------------------------- #
# integer frame arithmetic (must match bit-exactly)                           #
# --------------------------------------------------------------------------- #
def frame_count(samples, frame_length, hop_length):
    """stft.py:146-149 — frames WITHOUT torch.stft's centre padding."""
    return math.ceil(max(samples - frame_length, 0) / hop_length) + 1


def right_padding(samples, frame_length, hop_length):
    """stft.py:140-144 — zeros appended so the signal fills whole frames."""
    frames = frame_count(samples, frame_length, hop_length)
    return (frames - 1) * hop_length + frame_length - samples


def stft_frames(samples, frame_length, hop_length, n_fft=None, center=True):
    """Number of STFT frames torch.stft returns after stft.py:62 padding."""
    n_fft = frame_length if n_fft is None else n_fft
    padded = samples + right_padding(samples, frame_length, hop_length)
    if center:
        padded += 2 * (n_fft // 2)
    return 1 + (padded - n_fft) // hop_length


def istft_length(frames, hop_length, n_fft, center=True):
    """Samples torch.istft returns for `length=None`."""
    full = n_fft + hop_length * (frames - 1)
    return full - 2 * (n_fft // 2) if center else full


# --------------------------------------------------------------------------- #
# windows                                                                     #
# --------------------------------------------------------------------------- #
_COSINE_WINDOWS = {
    'hann': (0.5, 0.5),
    'hanning': (0.5, 0.5),
    'hamming': (0.54, 0.46),
    'blackman': (0.42, 0.5, 0.08),
}


def get_window(name, length):
    """Periodic ("fftbins=True") window as scipy.signal.get_window builds it.

    scipy builds a symmetric window of ``length + 1`` points and drops the last
    one; for the generalised-cosine family that is
    ``sum_k (-1)^k a_k cos(2 pi k n / length)``.  (stft.py:46-53)
    """
    if name is None or name in ('boxcar', 'rect', 'rectangular', 'ones'):
        return np.ones(length, dtype=np.float64)
    if name in _COSINE_WINDOWS:
        fac = np.linspace(-np.pi, np.pi, length + 1)
        w = np.zeros(length + 1)
        for k, a in enumerate(_COSINE_WINDOWS[name]):
            w += a * np.cos(k * fac)
        return w[:-1]
    import scipy.signal  # other names: defer to the same third-party call
    return scipy.signal.get_window(name, length)


def _padded_window(window, frame_length, n_fft):
    """torch.stft centres a short window inside n_fft (zero padded)."""
    window = np.asarray(window, dtype=np.float64)
    assert window.shape == (frame_length,)
    if n_fft == frame_length:
        return window
    left = (n_fft - frame_length) // 2
    out = np.zeros(n_fft)
    out[left:left + frame_length] = window
    return out


# --------------------------------------------------------------------------- #
# STFT / iSTFT                                                                #
# --------------------------------------------------------------------------- #
def stft(x, frame_length=512, hop_length=256, window='hann', normalized=True,
         onesided=True, compression_factor=1.0, scale_factor=1.0, n_fft=None,
         dtype=np.float64, center=True, pad_mode='constant', raw_framing=False):
    """STFT.forward, stft.py:59-89.

    ``X[k, t] = sum_n w~[n] x~[t H + n] exp(-2 pi i k n / N)``, x~ the signal
    padded by ``right_padding`` on the right (``STFT.pad``: ``F.pad(mode=pad_mode)``,
    stft.py:140-144; skipped with ``raw_framing``, i.e. plain ``torch.stft``) and, when
    ``center``, by ``N//2`` on both sides in ``pad_mode`` ('constant' zeros or 'reflect').
    """
    n_fft = frame_length if n_fft is None else n_fft
    x = np.asarray(x)
    lead = x.shape[:-1]
    samples = x.shape[-1]
    sig = x.reshape(-1, samples).astype(dtype)
    if isinstance(window, str) or window is None:
        window = get_window(window, frame_length)
    win = _padded_window(window, frame_length, n_fft).astype(dtype)
    pad_r = 0 if raw_framing else right_padding(samples, frame_length, hop_length)
    half = n_fft // 2 if center else 0
    mode = {'constant': 'constant', 'reflect': 'reflect'}[pad_mode]
    padded = np.pad(sig, ((0, 0), (0, pad_r)), mode=mode)
    padded = np.pad(padded, ((0, 0), (half, half)), mode=mode)
    n_frames = 1 + (padded.shape[1] - n_fft) // hop_length
    idx = (np.arange(n_frames)[:, None] * hop_length
           + np.arange(n_fft)[None, :])
    frames = padded[:, idx] * win  # (sig, T, N)
    if onesided:
        spec = np.fft.rfft(frames, axis=-1)
    else:
        spec = np.fft.fft(frames, axis=-1)
    spec = np.swapaxes(spec, -1, -2)  # (sig, F, T)
    if normalized:
        spec = spec / math.sqrt(float(np.sum(np.asarray(window, float) ** 2)))
    if compression_factor != 1:
        mag = np.abs(spec)
        spec = mag ** compression_factor * np.exp(1j * np.angle(spec))
    spec = spec * scale_factor
    return spec.reshape(*lead, *spec.shape[-2:])


def istft(spec, frame_length=512, hop_length=256, window='hann',
          normalized=True, onesided=True, compression_factor=1.0,
          scale_factor=1.0, n_fft=None):
    """STFT.backward, stft.py:101-138 (torch.istft, center=True, length=None).

    Per frame a real inverse DFT of the first ``N//2+1`` bins (imaginary parts
    of DC / Nyquist ignored, two-sided input sliced to the one-sided half),
    times the window, overlap-added at stride H, divided by the overlap-added
    squared window, trimmed by ``N//2`` at both ends.
    """
    n_fft = frame_length if n_fft is None else n_fft
    spec = np.asarray(spec).astype(np.complex128)
    lead = spec.shape[:-2]
    n_bins, n_frames = spec.shape[-2:]
    spec = spec.reshape(-1, n_bins, n_frames)
    if isinstance(window, str) or window is None:
        window = get_window(window, frame_length)
    win = _padded_window(window, frame_length, n_fft)
    spec = spec / scale_factor
    if compression_factor != 1:
        mag = np.abs(spec)
        spec = mag ** (1 / compression_factor) * np.exp(1j * np.angle(spec))
    if normalized:
        spec = spec * math.sqrt(float(np.sum(np.asarray(window, float) ** 2)))
    if not onesided:
        spec = spec[:, :n_fft // 2 + 1]
    assert spec.shape[1] == n_fft // 2 + 1
    frames = np.fft.irfft(np.swapaxes(spec, -1, -2), n=n_fft, axis=-1)
    frames = frames * win  # (sig, T, N)
    full = n_fft + hop_length * (n_frames - 1)
    out = np.zeros((spec.shape[0], full))
    env = np.zeros(full)
    for t in range(n_frames):
        out[:, t * hop_length:t * hop_length + n_fft] += frames[:, t]
        env[t * hop_length:t * hop_length + n_fft] += win ** 2
    half = n_fft // 2
    out = out[:, half:full - half]
    env = env[half:full - half]
    if env.size and np.abs(env).min() < 1e-11:
        raise RuntimeError('window overlap add min: 1')  # torch NOLA check
    out = out / env
    return out.reshape(*lead, -1)


def stft_grad(w, samples, frame_length=512, hop_length=256, window='hann', normalized=True,
              scale_factor=1.0, n_fft=None, center=True, pad_mode='constant', raw_framing=False):
    """Gradient of ``sum(Re X * Re w + Im X * Im w)`` w.r.t. the input of `stft` (no compression):
    the adjoint of the linear map, built by applying `stft` to the unit impulses' images --
    here simply via the real-linear structure: g = A^T w with A applied column by column
    would be O(S^2); instead overlap-add the inverse-DFT-like frames and fold the padding."""
    n_fft = frame_length if n_fft is None else n_fft
    w = np.asarray(w, dtype=np.complex128)
    lead = w.shape[:-2]
    W = w.reshape(-1, *w.shape[-2:])                       # (sig, F, T)
    if isinstance(window, str) or window is None:
        window = get_window(window, frame_length)
    win = _padded_window(window, frame_length, n_fft).astype(np.float64)
    scale = scale_factor / (math.sqrt(float(np.sum(np.asarray(window, float) ** 2))) if normalized else 1.0)
    pad_r = 0 if raw_framing else right_padding(samples, frame_length, hop_length)
    half = n_fft // 2 if center else 0
    total = samples + pad_r + 2 * half
    n_frames = W.shape[-1]
    # d/dframe of Re<rfft(frame), w> = Re(sum_k conj-weighted basis): irfft-like without 1/N
    k = np.arange(W.shape[-2])[:, None]
    n = np.arange(n_fft)[None, :]
    ang = 2 * np.pi * k * n / n_fft
    frames = (np.einsum('sft,fn->stn', W.real, np.cos(ang)) - np.einsum('sft,fn->stn', W.imag, np.sin(ang)))
    frames = frames * win * scale
    gp = np.zeros((W.shape[0], total))
    for t in range(n_frames):
        gp[:, t * hop_length:t * hop_length + n_fft] += frames[:, t]
    # fold the paddings back (adjoint of np.pad): centre first, then the right padding
    def unpad(g, left, right, length):
        core = g[:, left:left + length].copy()
        if pad_mode == 'reflect':
            for j in range(left):                          # padded[j] = x[left - j]
                core[:, left - j] += g[:, j]
            for j in range(right):                         # padded[left + length + j] = x[length - 2 - j]
                core[:, length - 2 - j] += g[:, left + length + j]
        return core
    g = unpad(gp, half, half, samples + pad_r)
    g = unpad(g, 0, pad_r, samples)
    return g.reshape(*lead, samples)


def manner_stft_loss(x, y, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50),
                     win_lengths=(600, 1200, 240), factor_sc=0.1, factor_mag=0.1):
    """MultiResolutionSTFTLoss, models/manner/stft_loss.py:22-151: per resolution raw
    ``torch.stft`` (reflect centre padding, periodic Hann of win_length inside fft_size, no
    normalisation), ``m = sqrt(max(re^2 + im^2, 1e-7))``, spectral convergence
    ``||m_y - m_x||_F / ||m_y||_F`` and ``mean |log m_y - log m_x|`` per item; mean over the
    resolutions, times the factors."""
    sc_tot, mag_tot = 0.0, 0.0
    for n_fft, hop, wl in zip(fft_sizes, hop_sizes, win_lengths):
        kw = dict(frame_length=wl, hop_length=hop, n_fft=n_fft, window='hann', normalized=False,
                  pad_mode='reflect', raw_framing=True)
        mags = []
        for sig in (x, y):
            spec = stft(sig, **kw)
            mags.append(np.sqrt(np.maximum(spec.real ** 2 + spec.imag ** 2, 1e-7)))
        mx, my = mags
        sc_tot = sc_tot + np.sqrt(((my - mx) ** 2).sum((-2, -1))) / np.sqrt((my ** 2).sum((-2, -1)))
        mag_tot = mag_tot + np.abs(np.log(my) - np.log(mx)).mean((-2, -1))
    return factor_sc * sc_tot / len(fft_sizes), factor_mag * mag_tot / len(fft_sizes)


def _conv_filters(frame_length, hop_length, window, normalized):
    """ConvSTFT.__init__, stft.py:213-238: rows k = 0..L/2 of the DFT matrix times the window
    (square root of the scipy window when given by name), DC row / sqrt(2), everything
    / (0.5 L / sqrt(H)) when normalised.  Returns the complex (F, L) bank and that factor."""
    if isinstance(window, str) or window is None:
        window = get_window(window, frame_length) ** 0.5
    window = np.asarray(window, dtype=np.float64)
    n = np.arange(frame_length)
    k = np.arange(frame_length // 2 + 1)[:, None]
    bank = np.exp(-2j * np.pi * ((k * n) % frame_length) / frame_length)
    bank[0] /= math.sqrt(2.0)
    factor = 0.5 * frame_length / math.sqrt(hop_length)
    if normalized:
        bank = bank / factor
    return bank * window, factor


def conv_stft(x, frame_length=512, hop_length=256, window='hann', compression_factor=1.0,
              scale_factor=1.0, normalized=True):
    """ConvSTFT.forward, stft.py:243-270: right pad to whole frames, L - H zeros on both
    sides (pad, stft.py:305-315), strided correlation with the filter bank, optional
    magnitude compression, scale."""
    x = np.asarray(x)
    lead, samples = x.shape[:-1], x.shape[-1]
    sig = x.reshape(-1, samples).astype(np.float64)
    bank, _ = _conv_filters(frame_length, hop_length, window, normalized)
    edge = frame_length - hop_length
    pad_r = right_padding(samples, frame_length, hop_length)
    padded = np.zeros((sig.shape[0], samples + pad_r + 2 * edge))
    padded[:, edge:edge + samples] = sig
    n_frames = (padded.shape[1] - frame_length) // hop_length + 1
    idx = np.arange(n_frames)[:, None] * hop_length + np.arange(frame_length)[None, :]
    spec = np.einsum('stn,kn->skt', padded[:, idx], bank)
    if compression_factor != 1:
        spec = np.abs(spec) ** compression_factor * np.exp(1j * np.angle(spec))
    spec = spec * scale_factor
    return spec.reshape(*lead, *spec.shape[-2:])


def conv_istft(spec, frame_length=512, hop_length=256, window='hann', compression_factor=1.0,
               scale_factor=1.0, normalized=True):
    """ConvSTFT.backward, stft.py:272-303: / scale, decompression, transposed convolution
    with the same real / imaginary filters (the adjoint of the analysis), / factor^2 when the
    filters are not normalised, L - H samples cut from both ends."""
    spec = np.asarray(spec).astype(np.complex128)
    lead = spec.shape[:-2]
    n_bins, n_frames = spec.shape[-2:]
    spec = spec.reshape(-1, n_bins, n_frames) / scale_factor
    if compression_factor != 1:
        spec = np.abs(spec) ** (1 / compression_factor) * np.exp(1j * np.angle(spec))
    bank, factor = _conv_filters(frame_length, hop_length, window, normalized)
    # real part of sum_k X[k, t] conj(bank[k, n]) = Re X * Re bank + Im X * Im bank
    frames = np.einsum('skt,kn->stn', spec.real, bank.real) + \
        np.einsum('skt,kn->stn', spec.imag, bank.imag)
    full = frame_length + hop_length * (n_frames - 1)
    out = np.zeros((spec.shape[0], full))
    for t in range(n_frames):
        out[:, t * hop_length:t * hop_length + frame_length] += frames[:, t]
    if not normalized:
        out = out / factor ** 2
    edge = frame_length - hop_length
    # ``x[..., padding:-padding]`` (stft.py:298): with hop == frame_length the slice is
    # ``[0:-0]`` = empty -- the reference returns no samples at all in that case (kept)
    out = out[:, edge:-edge] if edge else out[:, 0:0]
    return out.reshape(*lead, -1)


# --------------------------------------------------------------------------- #
# mel filterbank                                                              #
# --------------------------------------------------------------------------- #
def _linspace_f32(start, end, steps):
    """torch.linspace in float32: float32 step, symmetric halves around the
    midpoint, each point rounded once (fused multiply-add)."""
    start = np.float64(np.float32(start))
    end = np.float64(np.float32(end))
    step = np.float64(np.float32((end - start) / (steps - 1)))
    i = np.arange(steps)
    return np.where(i < steps // 2, start + step * i,
                    end - step * (steps - 1 - i)).astype(np.float32)


def fft_freqs(fs=16e3, n_fft=512):
    """utils.py:40-66, onesided."""
    freqs = np.arange(n_fft) * fs / n_fft
    return freqs[~(freqs > fs / 2)]


def mel_filterbank(n_filters=64, n_fft=512, fs=16e3, fmin=50, fmax=8000):
    """MelFilterbank.calc_filterbank, stft.py:161-176, in float32 like torch.

    Returns ``(filters (n_filters, F), fc (n_filters+2,), scaling (n_filters, 1))``.
    """
    mel_min = 2595 * math.log10(1 + fmin / 700)
    mel_max = 2595 * math.log10(1 + fmax / 700)
    mel = _linspace_f32(mel_min, mel_max, n_filters + 2)
    # torch evaluates 10**x with a ~1-ulp vectorised powf; a correctly rounded
    # power differs from it in isolated last bits, so these constants are
    # pinned to the reference within 1 ulp, not bit-exactly (the PRODUCT builds
    # them with the reference's own torch calls and is pinned bit-exactly).
    ratio = (mel / np.float32(2595)).astype(np.float32)
    power = np.power(10.0, ratio.astype(np.float64)).astype(np.float32)
    fc = (np.float32(700) * (power - np.float32(1))).astype(np.float32)
    f = fft_freqs(fs, n_fft).astype(np.float32)
    filters = np.zeros((n_filters, len(f)), dtype=np.float32)
    for row, i in enumerate(range(1, n_filters + 1)):
        up = (fc[i - 1] <= f) & (f <= fc[i])
        filters[row, up] = (f[up] - fc[i - 1]) / (fc[i] - fc[i - 1])
        down = (fc[i] <= f) & (f <= fc[i + 1])
        filters[row, down] = (fc[i + 1] - f[down]) / (fc[i + 1] - fc[i])
    scaling = filters.sum(axis=1, keepdims=True, dtype=np.float32)
    filters = filters / scaling
    return filters, fc, scaling


def mel_forward(filters, x):
    """MelFilterbank.forward, stft.py:189-190: filters @ x over (..., F, T)."""
    return np.matmul(filters.astype(np.float64), x)


def mel_backward(filters, scaling, x):
    """MelFilterbank.backward, stft.py:192-198: (filters*scaling).T @ x."""
    inv = (filters * scaling).T.astype(np.float64)
    return np.matmul(inv, x)


# --------------------------------------------------------------------------- #
# features, stacking, normalisation                                           #
# --------------------------------------------------------------------------- #
def fbe(spec, filters, normalize=False, compression='none'):
    """FeatureExtractor.fbe, features.py:186-200 (no DCT branch).

    ``spec``: ``(B, C, F, T)`` or ``(C, F, T)`` complex → ``(B, M, T)`` / ``(M, T)``.
    """
    spec = np.asarray(spec)
    unbatched = spec.ndim == 3
    if unbatched:
        spec = spec[None]
    if spec.ndim != 4:
        raise ValueError(f'input must be 3 or 4 dimensional, got {spec.ndim}')
    power = (np.abs(spec.astype(np.complex128)) ** 2).mean(1)
    out = mel_forward(filters, power)
    if normalize:
        out = out / (out.sum(1, keepdims=True) + EPS32)
    if compression == 'log':
        out = np.log(out + EPS32)
    elif compression == 'cubic':
        out = out ** (1 / 3)
    elif compression != 'none':
        raise ValueError('compression must be log, cubic or none, got '
                         f'{compression}')
    return out[0] if unbatched else out


def dct_features(energies, n_dct=14):
    """DCT branch of FeatureExtractor.fbe, features.py:199-219 (defaults: type 2, ortho,
    no DC term, deltas and double deltas): ``(B, M, T)`` -> ``(B, 3*(n_dct-1), T)``.

    scipy.fft.dct(type=2, norm='ortho') along the filter axis:
    ``y[k] = s_k * sum_n x[n] cos(pi k (2n+1) / (2M))``, ``s_0 = sqrt(1/M)``, ``s_k = sqrt(2/M)``;
    coefficients 1..n_dct-1 are kept (features.py:205-206).  Deltas are first / second
    differences along frames, zero-padded on the left (features.py:208-215).
    """
    x = np.asarray(energies, dtype=np.float64)
    m = x.shape[1]
    k = np.arange(1, n_dct)[:, None]
    n = np.arange(m)[None, :]
    basis = math.sqrt(2.0 / m) * np.cos(math.pi * k * (2 * n + 1) / (2 * m))
    cc = np.einsum('km,bmt->bkt', basis, x)
    d1 = np.zeros_like(cc)
    d1[..., 1:] = cc[..., 1:] - cc[..., :-1]
    d2 = np.zeros_like(cc)
    d2[..., 2:] = cc[..., 2:] - 2 * cc[..., 1:-1] + cc[..., :-2]
    return np.concatenate([cc, d1, d2], axis=1)


def _batched(spec):
    spec = np.asarray(spec).astype(np.complex128)
    unbatched = spec.ndim == 3
    if unbatched:
        spec = spec[None]
    if spec.ndim != 4:
        raise ValueError(f'input must be 3 or 4 dimensional, got {spec.ndim}')
    return spec, unbatched


def ild(spec, filters):
    """FeatureExtractor.ild, features.py:222-240: mel(20 log10((|R|+eps)/(|L|+eps)))."""
    spec, unbatched = _batched(spec)
    mag = np.abs(spec)
    out = mel_forward(filters, 20 * np.log10((mag[:, 1] + EPS32) / (mag[:, 0] + EPS32)))
    return out[0] if unbatched else out


def ipd(spec, filters):
    """FeatureExtractor.ipd, features.py:242-260: mel(angle(R) - angle(L)), not wrapped."""
    spec, unbatched = _batched(spec)
    phase = np.angle(spec)
    out = mel_forward(filters, phase[:, 1] - phase[:, 0])
    return out[0] if unbatched else out


def ic(spec, filters, hop_length=256, fs=16e3, tau=10e-3):
    """FeatureExtractor.ic, features.py:262-296.

    Auto / cross power spectra smoothed along frames by ``y[t] = (1-a) x[t] + a y[t-1]``,
    ``a = exp(-hop / (tau fs))`` (torchaudio.functional.lfilter with float32 coefficients,
    whose default ``clamp=True`` clips the *output* to [-1, 1]); coherence
    ``|phi_lr|^2 / (phi_ll phi_rr)``, mel projection, square root.
    """
    spec, unbatched = _batched(spec)
    alpha = math.exp(-hop_length / (tau * fs))
    a1 = float(np.float32(-alpha))             # a_coeffs = [1, -alpha], float32
    b0 = float(np.float32(1 - alpha))          # b_coeffs = [1 - alpha, 0]
    left, right = spec[:, 0], spec[:, 1]
    cross = left * np.conj(right)              # |L||R| exp(j(phase_L - phase_R))
    x = np.stack([np.abs(left) ** 2, np.abs(right) ** 2, cross.real, cross.imag])
    phi = np.zeros_like(x)
    prev = np.zeros(x.shape[:-1])
    for t in range(x.shape[-1]):
        prev = b0 * x[..., t] - a1 * prev
        phi[..., t] = prev
    phi = np.clip(phi, -1.0, 1.0)
    with np.errstate(divide='ignore', invalid='ignore'):
        coh = (phi[2] ** 2 + phi[3] ** 2) / (phi[0] * phi[1])
    out = np.sqrt(mel_forward(filters, coh))
    return out[0] if unbatched else out


DCT_FAMILY = {
    'mfcc': dict(compression='log'),
    'cubicmfcc': dict(compression='cubic'),
    'pdfcc': dict(normalize=True, compression='log'),
}
# rows the reference *declares* per feature (features.py:21-101); the DCT features declare 13
# but return 39 (deltas and double deltas are on by default) — quirk kept
DECLARED_ROWS = {'mfcc': 13, 'cubicmfcc': 13, 'pdfcc': 13}

FBE_FAMILY = {
    'fbe': dict(),
    'logfbe': dict(compression='log'),
    'cubicfbe': dict(compression='cubic'),
    'pdf': dict(normalize=True),
    'logpdf': dict(normalize=True, compression='log'),
    'cubicpdf': dict(normalize=True, compression='cubic'),
}


def extract_features(spec, filters, features, hop_length=256, fs=16e3):
    """FeatureExtractor.__call__, features.py:103-113: sorted names, cat on dim 0."""
    out, indices, start = [], {}, 0
    for name in sorted(features):
        if name in FBE_FAMILY:
            data = fbe(spec, filters, **FBE_FAMILY[name])
        elif name in DCT_FAMILY:
            unbatched = np.asarray(spec).ndim == 3
            e = fbe(spec, filters, **DCT_FAMILY[name])
            data = dct_features(e[None] if unbatched else e)
            data = data[0] if unbatched else data
        elif name in ('ild', 'ipd'):
            data = {'ild': ild, 'ipd': ipd}[name](spec, filters)
        elif name == 'ic':
            data = ic(spec, filters, hop_length=hop_length, fs=fs)
        else:
            raise ValueError(f'unrecognized feature, got {name}')
        out.append(data)
        indices[name] = (start, start + len(data))
        start += len(data)
    return np.concatenate(out, axis=0), indices


def stack(data, stacks):
    """FFNN.stack, ffnn.py:122-132: out[k*nf + f, t] = data[f, max(t-k, 0)]."""
    data = np.asarray(data)
    n_frames = data.shape[-1]
    out = [data]
    for k in range(1, stacks + 1):
        src = np.maximum(np.arange(n_frames) - k, 0)
        out.append(data[..., src])
    return np.concatenate(out, axis=0 if data.ndim == 2 else 1)


def decimate(data, decimation):
    """FFNN.decimate, ffnn.py:134-135."""
    return data[..., ::decimation]


def static_normalize(x, mean, std):
    """StaticNormalizer.forward, ffnn.py:186-187."""
    return (x - mean) / std


def cumulative_normalize(x, eps=1e-4):
    """CumulativeNormalizer.forward, ffnn.py:195-203."""
    x = np.asarray(x, dtype=np.float64)
    count = np.arange(1, x.shape[-1] + 1)
    mean = np.cumsum(x, -1) / count
    var = np.cumsum(x ** 2, -1) / count - mean ** 2
    return (x - mean) / np.sqrt(var + eps)


def training_statistics(items):
    """FFNN.pre_train, ffnn.py:137-148: per-utterance time means, averaged."""
    mean = sum(np.mean(x, -1, keepdims=True) for x in items) / len(items)
    msq = sum(np.mean(np.square(x), -1, keepdims=True) for x in items)
    var = msq / len(items) - mean ** 2
    return mean, np.sqrt(var)


def irm(foreground_mag, background_mag, filters):
    """FFNN.irm, ffnn.py:113-120 on (C, F, T) magnitudes."""
    fg = mel_forward(filters, (np.asarray(foreground_mag, float) ** 2).mean(0))
    bg = mel_forward(filters, (np.asarray(background_mag, float) ** 2).mean(0))
    return (1 + bg / (fg + EPS64)) ** -0.5


# --------------------------------------------------------------------------- #
# criteria                                                                    #
# --------------------------------------------------------------------------- #
def apply_mask(x, y, lengths):
    """criterion.py:229-234: zero everything at or beyond lengths[i]."""
    x, y = np.asarray(x), np.asarray(y)
    assert len(lengths) == x.shape[0]
    mask = np.zeros(x.shape)
    for i, length in enumerate(lengths):
        mask[i, ..., :int(length)] = 1
    return x * mask, y * mask


def snr(x, y, lengths):
    """criterion.py:75-101 → (B,) loss (negative SNR in dB)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    assert x.shape == y.shape and x.ndim >= 2
    x, y = apply_mask(x, y, lengths)
    ratio = (y ** 2).sum(-1) / (((y - x) ** 2).sum(-1) + EPS32)
    db = 10 * np.log10(ratio + EPS32)
    axes = tuple(range(1, x.ndim - 1))
    # quirk kept on purpose: for 2-D input `axes` is the empty tuple and
    # torch's mean(()) reduces over ALL dims -> a 0-dim batch mean
    # (this is what DCCRN hits: dccrn.py:115-122 passes (B, L) tensors)
    return -(db.mean(axes) if axes else db.mean())


def sisnr_matrix(x, y, lengths):
    """criterion.py:45-61 → (B, S_target, S_estimate) SI-SNR in dB."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    assert x.shape == y.shape and x.ndim == 3
    lengths = np.asarray(lengths)
    x, y = apply_mask(x, y, lengths)
    x = x - x.sum(2, keepdims=True) / lengths.reshape(-1, 1, 1)
    y = y - y.sum(2, keepdims=True) / lengths.reshape(-1, 1, 1)
    x, y = apply_mask(x, y, lengths)
    s_hat = x[:, None]            # (B, 1, S, L)
    s = y[:, :, None]             # (B, S, 1, L)
    s_target = (s_hat * s).sum(3, keepdims=True) * s \
        / (s ** 2).sum(3, keepdims=True)
    e_noise = s_hat - s_target
    ratio = (s_target ** 2).sum(3) / ((e_noise ** 2).sum(3) + EPS32)
    return 10 * np.log10(ratio + EPS32)


def sisnr(x, y, lengths):
    """criterion.py:21-72 → (B,) loss with PIT; also returns the arg-max perm."""
    mat = sisnr_matrix(x, y, lengths)
    n_src = mat.shape[1]
    perms = list(itertools.permutations(range(n_src)))
    # one_hot[p, i, perm[i]] = 1  (criterion.py:66-68)
    totals = np.stack([sum(mat[:, i, p[i]] for i in range(n_src))
                       for p in perms], axis=1)
    best = totals.argmax(1)
    return -totals.max(1) / n_src, np.asarray(perms)[best]


def snr_grad(x, y, lengths):
    """d snr(x, y, lengths).sum() / dx — closed form, SURVEY §8(a')."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    xm, ym = apply_mask(x, y, lengths)
    diff = ym - xm
    p = (ym ** 2).sum(-1, keepdims=True)
    d = (diff ** 2).sum(-1, keepdims=True)
    r = p / (d + EPS32)
    rows = int(np.prod(x.shape[1:-1])) if x.ndim > 2 else x.shape[0]
    coef = -(10 / math.log(10)) * 2 * p / ((r + EPS32) * (d + EPS32) ** 2)
    mask = apply_mask(np.ones_like(x), y, lengths)[0]
    return coef * diff * mask / rows


def sisnr_grad(x, y, lengths):
    """d sisnr(x, y, lengths).sum() / dx for the arg-max permutation (PIT)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    lengths = np.asarray(lengths)
    _, perm = sisnr(x, y, lengths)
    n_batch, n_src, _ = x.shape
    xm, ym = apply_mask(x, y, lengths)
    a = xm - xm.sum(2, keepdims=True) / lengths.reshape(-1, 1, 1)
    b = ym - ym.sum(2, keepdims=True) / lengths.reshape(-1, 1, 1)
    a, b = apply_mask(a, b, lengths)
    grad = np.zeros_like(x)
    k = 10 / math.log(10)
    for n in range(n_batch):
        for i in range(n_src):          # target i is matched with estimate j
            j = perm[n, i]
            aj, bi = a[n, j], b[n, i]
            dot, eb, ea = aj @ bi, bi @ bi, aj @ aj
            t = dot * dot / eb
            e = ea - t
            r = t / (e + EPS32)
            dt = 2 * dot / eb * bi
            de = 2 * aj - dt
            dr = (dt * (e + EPS32) - t * de) / (e + EPS32) ** 2
            g = -k / (r + EPS32) * dr / n_src
            g[int(lengths[n]):] = 0
            grad[n, j] += g
    return grad
