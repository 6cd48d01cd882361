"""The symmetry fold behind brv_stft_fold.cu, restated in numpy and checked against numpy's FFT
(no GPU): the N-point real DFT as four Q x Q contractions for N = 4Q and for N = 4Q - 2 (SGMSE's
510), and the transpose used by the inverse — including the placement of the four folded
segments on the hop blocks that the fused overlap-add relies on (hop = Q).

This is a model of OUR algorithm (index maps, rank-1 corrections, Hermitian weights), so that the
derivation in the kernel comments stays checkable without a B200; parity with the reference is
the job of the GPU tests and the oracle."""
import numpy as np
import pytest


def fold_forward(frame, N):
    """Windowed frame (N,) -> one-sided spectrum via four Q x Q contractions."""
    odd = N % 4 == 2
    Q = (N + 2) // 4
    Hf = N // 2
    n = np.arange(Q)
    a = frame[n].copy()
    b = frame[Hf - n].copy()
    c = frame[Hf + n].copy()
    d = np.where(n > 0, frame[(N - n) % N], 0.0)
    c[0] = 0.0                                   # n = 0: x[N/2] enters once (through b)
    ee, eo = a + d + b + c, a + d - b - c
    oe, oo = a - d - b + c, a - d + b - c
    m = np.arange(Q)[:, None]
    th_e = 2 * np.pi * (2 * m) * n[None, :] / N
    th_o = 2 * np.pi * (2 * m + 1) * n[None, :] / N
    re_e, re_o = np.cos(th_e) @ ee, np.cos(th_o) @ eo
    im_e, im_o = -np.sin(th_e) @ oe, -np.sin(th_o) @ oo
    F = N // 2 + 1
    X = np.zeros(F, dtype=complex)
    sg = (-1.0) ** np.arange(Q)
    if not odd:
        # the self-paired n = Q column (x[Q], x[3Q]) as rank-1 terms, and the Nyquist bin
        eeq, ooq = frame[Q] + frame[3 * Q], frame[Q] - frame[3 * Q]
        re_e = re_e + sg * eeq
        im_o = im_o - sg * ooq
        X[Hf] = np.sum(sg * ee) + eeq            # sum_n (-1)^n ee[n], Q even
    X[0:2 * Q:2] = re_e + 1j * im_e
    X[1:2 * Q:2] = re_o + 1j * im_o
    return X


def fold_inverse_segments(X, N, window):
    """One-sided spectrum -> windowed frame w * irfft(X), assembled the way the kernel's epilogue
    does it: four Q-long segments per frame, segment s landing on hop block s (hop = Q)."""
    odd = N % 4 == 2
    Q = (N + 2) // 4
    Hf = N // 2
    m = np.arange(Q)
    k_e, k_o = 2 * m, 2 * m + 1
    w_e = np.where(k_e == 0, 1.0, 2.0)
    w_o = np.where(2 * k_o == N, 1.0, 2.0)       # k = N/2 is an odd bin when N = 4Q - 2
    n = np.arange(Q)[:, None]
    Xe, Xo = X[0:2 * Q:2].copy(), X[1:2 * Q:2].copy()
    Xe[0] = Xe[0].real                           # Im X[0] ignored by the c2r inverse
    Ce = np.cos(2 * np.pi * k_e[None, :] * n / N) @ (w_e * Xe.real)
    Co = np.cos(2 * np.pi * k_o[None, :] * n / N) @ (w_o * Xo.real)
    Se = np.sin(2 * np.pi * k_e[None, :] * n / N) @ (w_e * Xe.imag)
    So = np.sin(2 * np.pi * k_o[None, :] * n / N) @ (w_o * Xo.imag)
    ny = 0.0 if odd else X[Hf].real
    sg = (-1.0) ** np.arange(Q)
    f0 = Ce + Co - Se - So + sg * ny             # f[n]
    f1 = Ce - Co + Se - So + sg * ny             # f[N/2 - n]
    f2 = Ce - Co - Se + So + sg * ny             # f[N/2 + n]
    f3 = Ce + Co + Se + So + sg * ny             # f[N - n]
    seg = np.zeros((4, Q))                       # seg[s, o] = frame position s Q + o
    o = np.arange(Q)
    if not odd:
        # pacc / racc of the builders: the n = Q column in fp32
        alt = (-1.0) ** m
        p_acc = np.sum(alt * w_e * Xe.real)
        r_acc = np.sum(alt * w_o * Xo.imag)
        seg[0] = f0
        seg[2] = f2
        seg[1, 1:] = f1[Q - o[1:]]
        seg[3, 1:] = f3[Q - o[1:]]
        seg[1, 0] = p_acc - r_acc + ny           # f[Q]
        seg[3, 0] = p_acc + r_acc + ny           # f[3Q]
    else:
        seg[0] = f0                              # positions n = o
        seg[1] = f1[Q - 1 - o]                   # N/2 - n = Q + o   ->  n = Q - 1 - o
        seg[2, :Q - 1] = f2[o[:Q - 1] + 1]       # N/2 + n = 2Q + o  ->  n = o + 1
        seg[2, Q - 1] = f3[Q - 1]                # position 3Q - 1 = N - (Q - 1)
        seg[3, :Q - 2] = f3[Q - 2 - o[:Q - 2]]   # N - n = 3Q + o    ->  n = Q - 2 - o
    frame = np.zeros(4 * Q)
    for s in range(4):
        frame[s * Q:(s + 1) * Q] = seg[s]
    return frame[:N] / N * window, frame[N:]


@pytest.mark.parametrize('N', [128, 256, 384, 512, 126, 254, 382, 510])
def test_fold_forward_equals_rfft(N):
    rng = np.random.default_rng(N)
    frame = rng.standard_normal(N) * np.hanning(N + 1)[:N]
    got = fold_forward(frame, N)
    ref = np.fft.rfft(frame)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize('N', [128, 256, 384, 512, 126, 254, 382, 510])
def test_fold_inverse_segments_equal_windowed_irfft(N):
    rng = np.random.default_rng(N + 1)
    F = N // 2 + 1
    X = rng.standard_normal(F) + 1j * rng.standard_normal(F)   # Im of DC / Nyquist must be ignored
    window = np.hanning(N + 1)[:N]
    got, spare = fold_inverse_segments(X, N, window)
    ref = np.fft.irfft(X, n=N) * window
    assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)
    assert np.all(spare == 0.0)                  # N = 4Q - 2: positions N, N + 1 receive nothing


def test_round_trip_through_both_folds():
    for N in (512, 510):
        rng = np.random.default_rng(7)
        x = rng.standard_normal(N)
        back, _ = fold_inverse_segments(fold_forward(x, N), N, np.ones(N))
        assert np.abs(back - x).max() < 1e-11
