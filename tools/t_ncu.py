"""Dev harness (GPU box): a few launches of the STFT / iSTFT pair for ncu to capture.
    ncu --set full -k regex:_t_kernel -s 4 -c 2 -o gpurun_out/t python tools/t_ncu.py [cfg2|cfg5]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brever_b200 as brv  # noqa: E402

CFG = {'cfg2': (dict(frame_length=512, hop_length=128), (64, 64000)),
       'cfg5': (dict(frame_length=256, hop_length=128, normalized=False), (1024, 64000)),
       'cfg4': (dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5,
                     scale_factor=0.15), (128, 128000))}

kw, shape = CFG[sys.argv[1] if len(sys.argv) > 1 else 'cfg2']
stft = brv.STFT(**kw)
x = 0.05 * torch.randn(*shape, device='cuda')
for _ in range(4):
    spec = stft(x)
    y = stft.backward(spec)
torch.cuda.synchronize()
