"""Dev harness (GPU box): device time of STFT.forward / STFT.backward alone for a few
transform configurations on the same batch, CUDA events around back-to-back launches
over rotating inputs (larger than L2 in total).

    python tools/stft_sweep.py [batch] [seconds] [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brever_b200 as brv  # noqa: E402

CASES = [
    dict(frame_length=512, hop_length=128),
    dict(frame_length=512, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=510, hop_length=128, normalized=False),
    dict(frame_length=510, hop_length=128, normalized=False, compression_factor=0.5, scale_factor=0.15),
    dict(frame_length=512, hop_length=256),
    dict(frame_length=256, hop_length=128, normalized=False),
]


def timed(fn, args_list, reps):
    for a in args_list:
        fn(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(args_list[i % len(args_list)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    dev = torch.device('cuda', 0)
    xs = [0.05 * torch.randn(batch, seconds * 16000, device=dev) for _ in range(3)]
    for kw in CASES:
        stft = brv.STFT(**kw)
        specs = [stft(x) for x in xs]
        fwd = timed(lambda x: stft(x), xs, reps)
        inv = timed(lambda s: stft.backward(s), specs, reps)
        frames = specs[0].shape[-1] * batch
        print(f'{kw}: forward {fwd:.1f} us ({fwd * 1e3 / frames:.2f} ns/frame)  '
              f'inverse {inv:.1f} us ({inv * 1e3 / frames:.2f} ns/frame)')


if __name__ == '__main__':
    main()
