"""Dev harness (GPU box): device time of single stages of a bench workload, each
timed alone through CUDA-graph replays (bench.time_stages).

    python tools/kbench.py cfg2 [stage,stage...] [reps]
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
    only = sys.argv[2].split(',') if len(sys.argv) > 2 and sys.argv[2] != 'all' else None
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    wl = bench.WORKLOADS[name]
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    pipe = bench.Pipeline(name, wl, dev)
    work, _ = bench.stage_work(wl)
    set_bytes = sum(work[s]['bytes'] for s in pipe.stage_names)
    n_sets = max(2, min(8, int(2.2 * 126e6 / max(set_bytes, 1)) + 1))
    sets = []
    for i in range(n_sets):
        mix, fg = bench.make_batch(wl, 1000 + i)
        sets.append((mix.to(dev), fg.to(dev)))
    ms = bench.time_stages(pipe, sets, reps, only)
    peaks = bench.load_peaks()
    tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('BRV_'))
    for stage, t in ms.items():
        gbs = work[stage]['bytes'] / (t * 1e-3) / 1e9
        print(f'{name} {stage}: {t * 1e3:.1f} us  {gbs:.0f} GB/s ({gbs / peaks["hbm"]:.1%} of HBM) {tag}')


if __name__ == '__main__':
    main()
