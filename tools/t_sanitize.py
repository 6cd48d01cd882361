"""Dev harness (GPU box): one small launch of every tcgen05 kernel variant (forward / inverse, both
layouts, gradients) for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/t_sanitize.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brever_b200 as brv  # noqa: E402
from brever_b200 import _lib  # noqa: E402

lib = _lib.lib()
torch.manual_seed(0)
for kw in (dict(frame_length=512, hop_length=128), dict(frame_length=256, hop_length=128, normalized=False),
           dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15)):
    stft = brv.STFT(**kw)
    x = 0.05 * torch.randn(5, 30000, device='cuda')
    for variant in (4, 2, 3, 5, 6):
        lib.brv_set_tc_variant(variant)
        xg = x.clone().requires_grad_(kw.get('compression_factor', 1.0) == 1.0)
        spec = stft(xg)
        y = stft.backward(spec)                                   # frame-major input
        y2 = stft.backward(spec.detach().contiguous())            # bin-major input
        if xg.requires_grad:
            y.square().sum().backward()
        torch.cuda.synchronize()
        print(kw, 'variant', variant, 'ok', float(y.abs().max()), float((y - y2).abs().max()), flush=True)
lib.brv_set_tc_variant(0)
