"""Dev harness (GPU box): per-phase %globaltimer stamps of CTA 0 of the folded STFT / iSTFT
kernels.  Needs a library built with -DBRV_PHASE_TIMING:

    NVCC_EXTRA=-DBRV_PHASE_TIMING python __graft_entry__.py --force
    python tools/phase_times.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brever_b200 as brv  # noqa: E402
from brever_b200 import _lib  # noqa: E402


def stamps():
    buf = (ctypes.c_ulonglong * 64)()
    fn = _lib.lib().brv_debug_phase_times
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    fn(buf, 64)
    return list(buf)


def main():
    dev = torch.device('cuda', 0)
    stft = brv.STFT(512, 128)
    x = 0.05 * torch.randn(64, 64000, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        spec = stft(x)
        y = stft.backward(spec)
    for rep in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        spec = stft(x)
        torch.cuda.synchronize()
        t = stamps()
        f = t[32:38]
        print('forward  (us since start): span %.1f  rowscale %.1f  build %.1f  mma-drain %.1f  epilogue %.1f'
              % tuple((f[i] - f[0]) / 1e3 for i in range(1, 6)))
        if '--chain' not in sys.argv:      # --chain: the spectrogram stays L2-resident, as in a step
            flush.zero_()
        torch.cuda.synchronize()
        y = stft.backward(spec)
        torch.cuda.synchronize()
        t = stamps()
        for n in range(2):
            s = t[n * 8:n * 8 + 8]
            print('inverse tile %d (us since tile-0 start): top %.1f  scouts-done %.1f  build %.1f  mma-drain %.1f  '
                  'epilogue %.1f  bar %.1f  copy-out+bar %.1f  (copy-out loop of warp 0 done %.1f)' % ((n,) + tuple((v - t[0]) / 1e3 for v in s)))


if __name__ == '__main__':
    main()
