"""Dev harness (GPU box): per-role %globaltimer stamps of CTA 0 of the transposed strip kernels
(brv_fold_t.cuh).  Needs a library built with -DBRV_PHASE_TIMING:

    NVCC_EXTRA=-DBRV_PHASE_TIMING python __graft_entry__.py --force
    python tools/t_phase.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brever_b200 as brv  # noqa: E402
from brever_b200 import _lib  # noqa: E402

ROLES = ['loader/scout', 'builder', 'mma', 'epilogue']


def stamps():
    buf = (ctypes.c_ulonglong * 160)()
    fn = _lib.lib().brv_debug_t_times
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    fn(buf, 160)
    return list(buf)


WAITS = {0: ('loader/scout', ['slot/span empty', 'scout bar', '', '']),
         1: ('builder', ['fwd: window+fold | inv: scale ready', 'fwd: split+store', 'stage empty', 'fwd: fence+arrive | inv: bar']),
         2: ('mma', ['tmem_empty', 'stage full', '', '']),
         3: ('epilogue', ['ri_full', 'tmem_full', 'epilogue bar', '']),
         4: ('tma', ['stage empty', '', '', ''])}


def waits():
    buf = (ctypes.c_ulonglong * 32)()
    fn = _lib.lib().brv_debug_t_waits
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    fn(buf, 32)
    out = []
    for r, (name, kinds) in WAITS.items():
        parts = [f'{k} {buf[r * 4 + i] / 1.9e3:.1f} us' for i, k in enumerate(kinds) if k]
        out.append(f'  waits {name:12s}: ' + ', '.join(parts))
    return '\n'.join(out) + '   (cycles / 1.9 GHz, whole kernel, CTA 0)'


def show(title, t, t_kernel_us):
    t0 = min(v for v in t if v > 0)
    print(f'--- {title}: kernel {t_kernel_us:.1f} us (events); us since the first stamp: '
          'wait-start / work-start / [mid] / done')
    for r, name in enumerate(ROLES):
        for n in range(8):
            s = t[(r * 8 + n) * 4:(r * 8 + n) * 4 + 4]
            if s[0] == 0 or s[0] < t0:
                continue

            def f(v):
                return '%6.1f' % ((v - t0) / 1e3) if v >= t0 else '   -  '
            print(f'  {name:12s} tile {n}: {f(s[0])} {f(s[1])} [{f(s[3])}] {f(s[2])}')


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


def main():
    dev = torch.device('cuda', 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for name, kw, shape in [('cfg2 512/128 64x4s', dict(frame_length=512, hop_length=128), (64, 64000)),
                            ('cfg5 256/128 512x4s', dict(frame_length=256, hop_length=128, normalized=False),
                             (512, 64000))]:
        stft = brv.STFT(**kw)
        x = 0.05 * torch.randn(*shape, device=dev)
        _lib.lib().brv_set_tc_variant(int(os.environ.get('T_PHASE_VARIANT', '0')))
        for _ in range(3):
            spec = stft(x)
            stft.backward(spec)
        flush.zero_()
        torch.cuda.synchronize()
        us = timed(lambda: stft(x))
        show(name + ' forward', stamps(), us)
        print(waits())
        flush.zero_()
        torch.cuda.synchronize()
        us = timed(lambda: stft.backward(spec))
        show(name + ' inverse', stamps(), us)
        print(waits())


if __name__ == '__main__':
    main()
