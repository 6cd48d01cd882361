#!/usr/bin/env python
"""Benchmark of the time-frequency front-end hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg1|cfg4|cfg5]

A "step" is one pass of the hot path over one batch of synthetic mixtures.  At
N=1 the default workload is BASELINE.json configs[1]: the DCCRN-style complex
STFT (512-pt, hop 128) -> iSTFT round trip + SI-SNR loss on a batch of 64 x 4 s
at 16 kHz.  With N > 1 every rank processes its own batch of that size (weak
scaling, utterances sharded by rank, no collective on the data path) and the
only NCCL traffic is the final all-reduce of the mean metric.

Prints ONE JSON line on rank 0 (see the contract in the task statement):
  value      : audio-seconds / second, inputs resident in HBM, CUDA-event timed
  e2e        : same metric through the public API with pinned HOST buffers,
               H2D + D2H copies inside the timed region
  roofline   : dominant kernel, algorithmic bytes or flops / measured duration
  cpu_baseline : the reference's CPU implementation (torch CPU port of its exact
               library calls, oracle/torch_port.py) timed on this host's cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

FS = 16000

WORKLOADS = {
    # name: (batch, channels, seconds, frame_length, hop, stft kwargs, description)
    'cfg1': dict(batch=16, channels=2, seconds=4, frame_length=512, hop=256, kw={},
                 desc='FFNN front-end: 16x2ch x 4 s, STFT 512/256 -> 64 log-mel -> stack 5 -> normalise; mean_c(X) -> iSTFT'),
    'cfg2': dict(batch=64, channels=1, seconds=4, frame_length=512, hop=128, kw={},
                 desc='DCCRN round trip: 64 x 4 s, STFT 512/128 -> iSTFT -> SI-SNR'),
    'cfg4': dict(batch=128, channels=1, seconds=8, frame_length=510, hop=128,
                 kw=dict(normalized=False, compression_factor=0.5, scale_factor=0.15),
                 desc='SGMSE+: 128 x 8 s, compressed STFT 510/128 (c=0.5, scale 0.15) -> iSTFT'),
    'cfg5': dict(batch=1024, channels=2, seconds=4, frame_length=256, hop=128,
                 kw=dict(normalized=False),
                 desc='TF-GridNet: 1024 x 2ch x 4 s, STFT 256/128 -> iSTFT of 1 source'),
}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], bf16=p['bf16_tflops'],
                    bf16_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0,
                source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, val in zip(names, r[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_batch(wl, seed):
    """Synthetic mixtures: mixture + foreground target, CPU generator."""
    from _util import synthetic_mixture
    shape = (wl['batch'], wl['channels'], wl['seconds'] * FS)
    if wl['channels'] == 1:
        shape = (wl['batch'], wl['seconds'] * FS)
    return synthetic_mixture(shape, seed)


# --------------------------------------------------------------------------- #
# algorithmic work                                               #
# --------------------------------------------------------------------------- #
def stage_work(wl):
    S = wl['seconds'] * FS
    N, H = wl['frame_length'], wl['hop']
    F = N // 2 + 1
    import math
    T = 1 + ((math.ceil(max(S - N, 0) / H)) * H + N + 2 * (N // 2) - N) // H
    n_in = wl['batch'] * wl['channels']
    n_out = wl['batch']  # one source / channel-mean goes back through the iSTFT
    out_len = H * (T - 1)
    work = {
        'stft': dict(bytes=4 * n_in * S + 8 * n_in * F * T, flops=2 * N * 2 * F * T * n_in),
        'istft': dict(bytes=8 * n_out * F * T + 4 * n_out * out_len,
                      flops=2 * N * 2 * F * T * n_out),
        'sisnr': dict(bytes=2 * 4 * n_out * S, flops=0),
        'features': dict(bytes=8 * n_in * F * T + 4 * wl['batch'] * 384 * T, flops=0),
    }
    return work, T


# --------------------------------------------------------------------------- #
# our arm                                                                     #
# --------------------------------------------------------------------------- #
class Pipeline:
    """The hot path through the public drop-in API (what a brever model calls)."""

    def __init__(self, name, wl, device):
        import brever_b200 as brv
        self.brv, self.name, self.wl, self.device = brv, name, wl, device
        self.stft = brv.STFT(frame_length=wl['frame_length'], hop_length=wl['hop'], **wl['kw'])
        self.samples = wl['seconds'] * FS
        self.lengths = torch.full((wl['batch'],), self.samples, dtype=torch.int64, device=device)
        if name == 'cfg1':
            self.front = brv.ffnn.FFNNFrontEnd()
            self.front.stft = self.stft
            self.mean = torch.zeros(384, 1, device=device)
            self.std = torch.ones(384, 1, device=device)
        self.stage_names = {'cfg1': ['stft', 'features', 'istft', 'sisnr'],
                            'cfg2': ['stft', 'istft', 'sisnr'],
                            'cfg4': ['stft', 'istft', 'sisnr'],
                            'cfg5': ['stft', 'istft', 'sisnr']}[name]

    def step(self, mix, target, marks=None):
        """mix/target: device tensors.  Returns the (batch,) loss tensor."""
        def mark(i):
            if marks is not None:
                marks[i].record()
        mark(0)
        spec = self.stft(mix)
        mark(1)
        k = 2
        if self.name == 'cfg1':
            feats = self.front.features(spec, self.mean, self.std)   # noqa: F841
            mark(k)
            k += 1
            spec = spec.mean(1)
        elif self.name == 'cfg5':
            spec = spec[:, 0]
        y = self.stft.backward(spec)[..., :self.samples]
        mark(k)
        tgt = target if target.ndim == 2 else target.mean(1)
        loss = self.brv.sisnr(y.unsqueeze(1), tgt.unsqueeze(1), self.lengths)
        mark(k + 1)
        return loss


def run_ours(args, rank, world, device):
    import torch.distributed as dist
    from brever_b200 import _lib
    wl = WORKLOADS[args.workload]
    peaks = load_peaks()
    pipe = Pipeline(args.workload, wl, device)
    work, n_frames = stage_work(wl)
    audio_s = wl['batch'] * wl['seconds']            # per rank per step

    # rotating input sets so that a step never finds its inputs in L2
    set_bytes = sum(work[s]['bytes'] for s in pipe.stage_names)
    n_sets = max(2, min(8, int(2.2 * 126e6 / max(set_bytes, 1)) + 1))
    sets = []
    for i in range(n_sets):
        mix, fg = make_batch(wl, 1000 + rank * 16 + i)
        sets.append((mix.to(device), fg.to(device)))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_marks = len(pipe.stage_names) + 1
    for i in range(args.warmup):
        pipe.step(*sets[i % n_sets])
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(n_marks)]
             for _ in range(args.steps)]
    sampler = ClockSampler(device.index)
    barrier()
    sampler.start()
    launches0 = _lib.lib().brv_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    total = torch.zeros((), device=device)
    for i in range(args.steps):
        loss = pipe.step(*sets[i % n_sets], marks=marks[i])
        total += loss.mean()
    if world > 1:   # the one collective: final metric all-reduce (training.py:369-373)
        dist.all_reduce(total)
    stop.record()
    barrier()
    launches = _lib.lib().brv_launch_count() - launches0
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    t = torch.tensor([elapsed_ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t)
    ms_per_step = elapsed_ms / args.steps
    value = world * audio_s / (ms_per_step * 1e-3)

    # per-stage device time (events inside the timed region, same stream)
    stage_ms = {}
    for j, name in enumerate(pipe.stage_names):
        stage_ms[name] = statistics.mean(m[j].elapsed_time(m[j + 1]) for m in marks)
    dominant = max(stage_ms, key=stage_ms.get)
    w = work[dominant]
    if w['flops'] > 0:
        # DFT contraction: tensor-bound at peak/3 (3-product split precision)
        achieved = w['flops'] / (stage_ms[dominant] * 1e-3) / 1e12
        peak = peaks['bf16'] / 3
        roofline = dict(bound='tensor', kernel=dominant, achieved=round(achieved, 3),
                        peak=round(peak, 1), unit='TFLOP/s', frac=round(achieved / peak, 4),
                        traffic=None,
                        peak_source=peaks['source'] + ': bf16 dense / 3 (fp32-grade split-precision DFT GEMM, SURVEY 8d)',
                        hbm_frac=round(w['bytes'] / (stage_ms[dominant] * 1e-3) / 1e9 / peaks['hbm'], 4))
    else:
        achieved = w['bytes'] / (stage_ms[dominant] * 1e-3) / 1e9
        roofline = dict(bound='hbm', kernel=dominant, achieved=round(achieved, 1),
                        peak=peaks['hbm'], unit='GB/s', frac=round(achieved / peaks['hbm'], 4),
                        traffic=None, peak_source=peaks['source'])
    roofline['stage_ms'] = {k: round(v, 4) for k, v in stage_ms.items()}
    roofline['stage_hbm_frac'] = {
        k: round(work[k]['bytes'] / (v * 1e-3) / 1e9 / peaks['hbm'], 4) for k, v in stage_ms.items()}

    # ---- end to end: pinned host buffers, H2D + D2H inside the timed region ----
    host = [(m.cpu().pin_memory(), f.cpu().pin_memory()) for m, f in sets[:2]]
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 4
    d2h = wl['batch'] * 4
    copy_stream = torch.cuda.Stream(device)
    dev_bufs = [(torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    result = torch.empty(wl['batch'], dtype=torch.float32).pin_memory()

    def e2e_loop(steps):
        main = torch.cuda.current_stream(device)
        for i in range(steps + 1):
            if i < steps:                  # stage the next step's inputs
                b = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[b])
                    dev_bufs[b][0].copy_(host[i % 2][0], non_blocking=True)
                    dev_bufs[b][1].copy_(host[i % 2][1], non_blocking=True)
                    ready[b].record(copy_stream)
            if i >= 1:                     # compute step i-1
                b = (i - 1) % 2
                main.wait_event(ready[b])
                loss = pipe.step(*dev_bufs[b])
                consumed[b].record(main)
                result.copy_(loss, non_blocking=True)   # D2H read of the step result
        main.synchronize()

    for ev in consumed:
        ev.record(torch.cuda.current_stream(device))
    e2e_loop(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * audio_s * args.steps / float(t)

    out = {
        'metric': 'audio-seconds/sec (STFT->iSTFT->SI-SNR front-end)',
        'value': round(value, 1), 'unit': 'audio-s/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms_per_step, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"{args.workload}: {wl['desc']}",
                   'per_gpu_batch': wl['batch'], 'seconds': wl['seconds'], 'fs': FS,
                   'frame_length': wl['frame_length'], 'hop_length': wl['hop'],
                   'frames': n_frames, 'parallelism': f'dp{world} (utterances sharded by rank)',
                   'l2': f'{n_sets} rotating input sets (> 2x L2) so steps never hit L2-resident inputs',
                   'stft_path': os.environ.get('BRV_FORCE_GENERIC', '0') == '1' and 'generic' or 'default'},
        'e2e': {'value': round(e2e_value, 1), 'unit': 'audio-s/s',
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'note': 'pinned host -> device copy of mixture+target every step, double-buffered on a copy stream; loss read back every step'},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'clocks': clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_reference(args.workload, budget_s=15.0)
    return out


# --------------------------------------------------------------------------- #
# reference arm: the reference's CPU implementation on this host's cores      #
# --------------------------------------------------------------------------- #
def cpu_reference(workload, budget_s=15.0, steps=None, warmup=1):
    """Times oracle/torch_port.py (the reference's exact torch calls, float32 CPU,
    all host threads) on a bounded slice of the workload."""
    from oracle import tf_oracle as O
    from oracle import torch_port as P
    wl = WORKLOADS[workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    win = torch.from_numpy(O.get_window('hann', wl['frame_length']))
    kw = dict(frame_length=wl['frame_length'], hop_length=wl['hop'], **wl['kw'])
    filters = torch.from_numpy(O.mel_filterbank()[0]) if workload == 'cfg1' else None
    S = wl['seconds'] * FS

    def run(mix, fg):
        with torch.no_grad():
            spec = P.stft(mix, win, **kw)
            if workload == 'cfg1':
                feats = P.stack(P.logfbe(spec, filters), 5)
                feats = P.static_normalize(feats, 0.0, 1.0)  # noqa: F841
                spec = spec.mean(1)
            elif workload == 'cfg5':
                spec = spec[:, 0]
            y = P.istft(spec, win, **kw)[..., :S]
            tgt = fg if fg.ndim == 2 else fg.mean(1)
            lengths = torch.full((mix.shape[0],), S)
            return P.sisnr(y.unsqueeze(1), tgt.unsqueeze(1), lengths)

    # bounded sample: shrink the batch until one pass fits the budget
    sample = wl['batch']
    small = dict(wl, batch=min(4, wl['batch']))
    mix, fg = make_batch(small, 999)
    run(mix, fg)
    t0 = time.perf_counter()
    run(mix, fg)
    per_item = (time.perf_counter() - t0) / small['batch']
    reps = steps or 5
    while sample > 1 and per_item * sample * (reps + warmup) > budget_s:
        sample //= 2
    mix, fg = make_batch(dict(wl, batch=sample), 1000)
    for _ in range(warmup):
        run(mix, fg)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run(mix, fg)
        times.append(time.perf_counter() - t0)
    best, mean = min(times), statistics.mean(times)
    audio = sample * wl['seconds']
    return {'value': round(audio / mean, 1), 'unit': 'audio-s/s', 'cores': threads,
            'kind': 'port',
            'sample': f'{sample} of {wl["batch"]} utterances x {wl["seconds"]} s, {reps} passes '
                      f'(mean {mean * 1e3:.1f} ms, best {best * 1e3:.1f} ms), torch {torch.__version__} CPU float32',
            'best_value': round(audio / best, 1), 'ms_per_pass': round(mean * 1e3, 2),
            'sample_batch': sample}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    base = cpu_reference(args.workload, budget_s=60.0, steps=args.steps, warmup=args.warmup)
    wl = WORKLOADS[args.workload]
    return {
        'impl': 'reference',
        'metric': 'audio-seconds/sec (STFT->iSTFT->SI-SNR front-end)',
        'value': base['value'], 'unit': 'audio-s/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': base['ms_per_pass'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"{args.workload}: {wl['desc']}", 'per_gpu_batch': wl['batch'],
                   'seconds': wl['seconds'], 'fs': FS, 'frame_length': wl['frame_length'],
                   'hop_length': wl['hop'],
                   'note': 'reference CPU path (torch.stft/istft + criterion, brever call sequence) on host cores; each step is a bounded batch slice'},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: brever_b200 has no CPU path '
                         '(use --impl reference for the CPU baseline)')
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    try:
        out = run_ours(args, rank, world, device)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
