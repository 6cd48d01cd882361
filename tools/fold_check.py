"""Dev harness (GPU box): folded STFT kernels vs the dense tensor-core path, the
generic path and the float64 oracle, plus CUDA-event timings per variant."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import brever_b200 as brv  # noqa: E402
from brever_b200 import _lib  # noqa: E402
from oracle import tf_oracle as O  # noqa: E402
from _util import crandn, randn, rel_err, synthetic_mixture  # noqa: E402

DEV = 'cuda'
lib = _lib.lib()


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def check_forward():
    cases = [dict(frame_length=512, hop_length=128), dict(frame_length=512, hop_length=256),
             dict(frame_length=256, hop_length=128, normalized=False),
             dict(frame_length=400, hop_length=100, n_fft=512),
             dict(frame_length=512, hop_length=100, window='hamming'),
             dict(frame_length=128, hop_length=32),
             dict(frame_length=512, hop_length=128, compression_factor=0.5, scale_factor=0.15),
             dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5,
                  scale_factor=0.15)]
    ok = True
    for kw in cases:
        for samples in (100, 4097, 40000):
            x = randn((3, samples), 77)
            x[1] *= 1e-3
            x[2] *= 300.0
            stft = brv.STFT(**kw)
            new = stft(x.to(DEV)).cpu().numpy()
            ref = O.stft(x.numpy(), **kw)
            errs = [rel_err(new[i], ref[i]) for i in range(3)]
            worst = max(e[0] for e in errs)
            flag = 'ok ' if worst < 1e-4 else 'BAD'
            ok &= worst < 1e-4
            print(f'[fwd] {flag} {kw} S={samples}: {worst:.2e}', flush=True)
    return ok


def check_inverse():
    cases = [dict(frame_length=512, hop_length=128), dict(frame_length=512, hop_length=256),
             dict(frame_length=256, hop_length=128, normalized=False),
             dict(frame_length=400, hop_length=100, n_fft=512),
             dict(frame_length=128, hop_length=32),
             dict(frame_length=512, hop_length=128, compression_factor=0.5, scale_factor=0.15),
             dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5,
                  scale_factor=0.15)]
    ok = True
    for kw in cases:
        for frames in (1, 9, 200, 501):
            for layout in ('bin_major', 'frame_major'):
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
                    continue
                new = stft.backward(dev).cpu().numpy()
                worst = max(rel_err(new[i], ref[i])[0] for i in range(3))
                flag = 'ok ' if worst < 1e-4 else 'BAD'
                ok &= worst < 1e-4
                print(f'[inv] {flag} {kw} T={frames} {layout}: {worst:.2e}', flush=True)
    return ok


def bench():
    # variant 0: transposed strip kernels; 4: one-tile-per-TMEM kernels; 1: dense contraction
    variants = [int(v) for v in os.environ.get("FOLD_CHECK_VARIANTS", "0,4").split(",")]
    for name, shape, kw in [('cfg2', (64, 64000), dict(frame_length=512, hop_length=128)),
                            ('cfg1', (32, 64000), dict(frame_length=512, hop_length=256)),
                            ('cfg4', (128, 128000), dict(frame_length=510, hop_length=128, normalized=False,
                                                         compression_factor=0.5, scale_factor=0.15)),
                            ('cfg5', (2048, 64000), dict(frame_length=256, hop_length=128,
                                                         normalized=False))]:
        mix, _ = synthetic_mixture(shape, 1000)
        x = mix.to(DEV)
        stft = brv.STFT(**kw)
        spec = stft(x)
        spec_bm = spec.contiguous()          # bin-major copy (what a network hands to the iSTFT)
        for variant in variants:
            lib.brv_set_tc_variant(variant)
            t_f = timed(lambda: stft(x))
            t_i = timed(lambda: stft.backward(spec))
            t_b = timed(lambda: stft.backward(spec_bm))
            print(f'[time] {name} variant={variant}: stft {t_f:.1f} us, istft {t_i:.1f} us, '
                  f'istft(bin-major) {t_b:.1f} us', flush=True)
        lib.brv_set_tc_variant(0)
        del x, spec, spec_bm


if __name__ == '__main__':
    what = sys.argv[1:] or ['fwd', 'inv', 'bench']
    ok = True
    if 'fwd' in what:
        ok &= check_forward()
    if 'inv' in what:
        ok &= check_inverse()
    if 'bench' in what:
        bench()
    print('ALL OK' if ok else 'FAILURES')
