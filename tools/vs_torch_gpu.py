"""Dev harness (GPU box): the cfg2 round trip (STFT 512/128 -> iSTFT -> SI-SNR, 64 x 4 s)
through brever_b200 versus the same torch library calls the reference makes
(torch.stft / torch.istft -> cuFFT, elementwise criterion) on the same B200 — the
"existing Blackwell library" bar of SURVEY 8d.  Forward only and forward + backward,
eager launches, CUDA-event timed.

    python tools/vs_torch_gpu.py [iters]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
import brever_b200 as brv  # noqa: E402

EPS = torch.finfo(torch.float32).eps


def torch_round_trip(x, tgt, window, lengths):
    """stft.py:59-89,101-138 + criterion.py:41-72 with S = 1, on whatever device x is on."""
    norm = window.pow(2).sum().sqrt()
    spec = torch.stft(x, n_fft=512, hop_length=128, win_length=512, window=window, center=True,
                      pad_mode='constant', normalized=False, onesided=True, return_complex=True)
    spec = spec / norm
    y = torch.istft(spec * norm, n_fft=512, hop_length=128, win_length=512, window=window,
                    center=True, normalized=False, onesided=True, return_complex=False)
    y = y[..., :x.shape[-1]]
    mask = (torch.arange(x.shape[-1], device=x.device)[None] < lengths[:, None]).float()
    a, b = y * mask, tgt * mask
    a = (a - a.sum(-1, keepdim=True) / lengths[:, None]) * mask
    b = (b - b.sum(-1, keepdim=True) / lengths[:, None]) * mask
    proj = (a * b).sum(-1, keepdim=True) * b / b.pow(2).sum(-1, keepdim=True)
    noise = a - proj
    return -10 * torch.log10(proj.pow(2).sum(-1) / (noise.pow(2).sum(-1) + EPS) + EPS)


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    dev = torch.device('cuda', 0)
    wl = bench.WORKLOADS['cfg2']
    mix, fg = bench.make_batch(wl, 1000)
    mix, fg = mix.to(dev), fg.to(dev)
    lengths = torch.full((wl['batch'],), mix.shape[-1], dtype=torch.int64, device=dev)
    stft = brv.STFT(512, 128)
    window = stft.window.float().to(dev)

    def ours(grad):
        x = mix.clone().requires_grad_(True) if grad else mix
        loss = brv.sisnr(stft.backward(stft(x))[..., :mix.shape[-1]].unsqueeze(1),
                         fg.unsqueeze(1), lengths)
        if grad:
            loss.sum().backward()
        return loss

    def lib(grad):
        x = mix.clone().requires_grad_(True) if grad else mix
        loss = torch_round_trip(x, fg, window, lengths)
        if grad:
            loss.sum().backward()
        return loss

    a, b = ours(False), lib(False)
    print(f'max |loss difference| ours vs torch-GPU: {float((a - b).abs().max()):.2e} dB')
    audio = wl['batch'] * wl['seconds']
    for name, fn in (('brever_b200', ours), ('torch (cuFFT) on the same GPU', lib)):
        f = timed(lambda: fn(False), iters)
        fb = timed(lambda: fn(True), iters)
        print(f'{name:32s} forward {f * 1e3:8.1f} us ({audio / f * 1e3:10.0f} audio-s/s)   '
              f'forward+backward {fb * 1e3:8.1f} us')


if __name__ == '__main__':
    main()
