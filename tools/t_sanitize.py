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
    for variant in (4, 2, 3, 5, 6, 7, 0):
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
# float64 direct-sum kernels, ConvSTFT on them, the feature kernel, the small round-2 kernels
stft = brv.STFT(512, 128)
x64 = torch.randn(2, 5000, device='cuda', dtype=torch.float64, requires_grad=True)
y64 = stft.backward(stft(x64))
y64.square().sum().backward()
conv = brv.ConvSTFT(frame_length=400, hop_length=100)
xc = torch.randn(2, 5000, device='cuda', requires_grad=True)
conv.backward(conv(xc)).square().sum().backward()
front = brv.ffnn.FFNNFrontEnd()
mix = 0.05 * torch.randn(3, 2, 16000, device='cuda')
spec = front.stft(mix)
feats = front.features(spec)
m = brv.ffnn.channel_mean(spec)
tot = torch.zeros((), device='cuda')
brv.ffnn.accumulate_mean(tot, torch.randn(7, device='cuda'))
a, b = brv.modules.specfmt.split(spec, 'mag_phase')
brv.modules.specfmt.join(a, b, 'mag_phase')
torch.cuda.synchronize()
print('float64 / conv / features / specfmt ok', float(y64.abs().max()), tuple(feats.shape), flush=True)
