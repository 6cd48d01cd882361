"""Dev harness (GPU box): iSTFT kernel variants over launch sizes, to place the dispatch thresholds.
variant 0: default dispatch; 4: one-tile-per-TMEM kernel; 6 / 7: strip kernel, 64- / 32-frame tiles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brever_b200 as brv  # noqa: E402
from brever_b200 import _lib  # noqa: E402
from tools.fold_check import timed  # noqa: E402

lib = _lib.lib()
variants = [int(v) for v in os.environ.get('FOLD_CHECK_VARIANTS', '4,6,7').split(',')]
for kw in (dict(frame_length=512, hop_length=128), dict(frame_length=512, hop_length=256),
           dict(frame_length=256, hop_length=128), dict(frame_length=256, hop_length=64)):
    for n_sig in (8, 32, 64, 128, 256, 1024):
        stft = brv.STFT(**kw)
        x = 0.05 * torch.randn(n_sig, 64000, device='cuda')
        lib.brv_set_tc_variant(4)
        spec = stft(x)
        spec_bm = spec.contiguous()
        out = []
        for v in variants:
            lib.brv_set_tc_variant(v)
            out.append('v%d %7.1f / %7.1f' % (v, timed(lambda: stft.backward(spec)), timed(lambda: stft.backward(spec_bm))))
        lib.brv_set_tc_variant(0)
        print(f"{kw['frame_length']}/{kw['hop_length']} n_sig {n_sig:5d} cols/SM {n_sig * (64000 // kw['hop_length'] + 1) // 148:6d}: "
              + '   '.join(out) + '   (frame-major / bin-major us)', flush=True)
