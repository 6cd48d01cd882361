"""Dev harness (GPU box): run the stages of one bench workload a few times with
rotating inputs — the short command ncu wraps (tools/profile.sh) — and print
CUDA-event timings per stage when not under a profiler.

    python tools/stage_run.py [cfg2] [iters]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    wl = bench.WORKLOADS[name]
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    pipe = bench.Pipeline(name, wl, dev)
    sets = []
    for i in range(3):
        mix, fg = bench.make_batch(wl, 1000 + i)
        sets.append((mix.to(dev), fg.to(dev)))
    n_marks = len(pipe.stage_names) + 1
    for i in range(3):
        pipe.step(*sets[i % 3])
    torch.cuda.synchronize()
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(n_marks)] for _ in range(iters)]
    for i in range(iters):
        pipe.step(*sets[i % 3], marks=marks[i])
    torch.cuda.synchronize()
    for j, stage in enumerate(pipe.stage_names):
        ts = sorted(m[j].elapsed_time(m[j + 1]) for m in marks)
        print(f'{name} {stage}: median {ts[len(ts) // 2] * 1e3:.1f} us, min {ts[0] * 1e3:.1f} us')


if __name__ == '__main__':
    main()
