"""The tcgen05 STFT path against the generic CUDA-core path, the float64 oracle
and the golden vectors, with the measured error printed (pytest -s / log)."""
import numpy as np
import pytest
import torch

import brever_b200 as brv
from brever_b200 import _lib
from oracle import tf_oracle as O

from _util import assert_parity, golden, randn, rel_err, synthetic_mixture

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def cpu(t):
    return t.detach().cpu().numpy()


class generic_path:
    def __enter__(self):
        self.prev = _lib.lib().brv_set_force_generic(1)

    def __exit__(self, *a):
        _lib.lib().brv_set_force_generic(self.prev)


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
