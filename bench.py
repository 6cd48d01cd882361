#!/usr/bin/env python
"""Benchmark of the time-frequency front-end hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg1|cfg3|cfg4|cfg5]

A "step" is one pass of the hot path over one batch of synthetic mixtures.  At
N=1 the default workload is BASELINE.json configs[1]: the DCCRN-style complex
STFT (512-pt, hop 128) -> iSTFT round trip + SI-SNR loss on a batch of 64 x 4 s
at 16 kHz.  With N > 1 every rank processes its own batch of that size (weak
scaling, utterances sharded by rank, no collective on the data path) and the
only NCCL traffic is the final all-reduce of the mean metric; `--scaling strong`
shards the workload's fixed global batch instead.

Besides the headline the same JSON line carries (skip with --no-extras):
  workloads  : (N = 1) all five BASELINE.json configs measured in this run
  torch_gpu_baseline : (N = 1) the reference's torch.stft / istft calls on the same GPU
  strong_scaling : (N > 1) cfg5 (1024 x 2ch x 4 s) and cfg3 (256 x 4 s) at their fixed
               global batch, sharded with brever_b200.distributed.shard_bounds
  cpu_baseline_1thread : the CPU reference on one core

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
    'cfg3': dict(batch=256, channels=1, seconds=4, frame_length=512, hop=128, kw={},
                 desc='Conv-TasNet-style criterion: SI-SNR on 256 x 4 s resynthesised waveforms (no STFT on the path)'),
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


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index`, so that the pinned
    host buffers allocated afterwards are first touched on the GPU's NUMA node (eight ranks
    streaming H2D from one node's memory is what stopped the end-to-end number scaling in
    round 1).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        if not uuid.startswith('GPU-'):
            uuid = 'GPU-' + uuid
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 'nvml reported no local cpus inside the cpuset of this process'
        os.sched_setaffinity(0, cpus)
        return f'process bound to the {len(cpus)} cpus local to the gpu (nvml)'
    except Exception as exc:        # noqa: BLE001 -- best effort, never fails the bench
        return f'not bound ({type(exc).__name__})'


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region.

    NVML in-process (pynvml, ~1 ms period: the timed region of a 50-step run is only
    a few milliseconds long); `nvidia-smi -lms` as a fallback, started early because
    its start-up stalls CUDA launches of other processes for ~100 ms."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap', 0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index):
        self.index, self.rows, self.handle, self.proc = index, [], None, None
        self.stop_flag = threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                if not uuid.startswith('GPU-'):
                    uuid = 'GPU-' + uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                t = time.perf_counter()
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                mask = get_reasons(self.handle)
                self.rows.append((t, float(mhz), int(mask)))
            except Exception:
                pass
            time.sleep(0.0005)

    def _read_smi(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line))

    def start(self):
        """Call BEFORE the warm-up so start-up costs stay out of the timed region."""
        if self.handle is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap')
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}',
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
            t_end = time.time() + 5.0
            while not self.rows and time.time() < t_end:   # wait out nvidia-smi's start-up
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (perf_counter; all if None)."""
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
        if self.handle is None and self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no NVML / nvidia-smi'], 'samples': 0}
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        window = 'timed region'
        if not rows:                       # region shorter than one sampling period
            rows, window = self.rows[-3:], 'nearest samples (region shorter than the sampling period)'
        sm, reasons, mx = [], set(), self.max_mhz
        if self.handle is not None:
            for _, mhz, mask in rows:
                sm.append(mhz)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        reasons.add(name)
        else:
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            for _, line in rows:
                c = [v.strip() for v in line.split(',')]
                try:
                    sm.append(float(c[0]))
                    mx = float(c[1])
                    reasons.update(n for n, v in zip(names, c[2:6]) if v.lower().startswith('active'))
                except (ValueError, IndexError):
                    continue
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window,
                'source': 'nvml' if self.handle is not None else 'nvidia-smi'}


def make_batch(wl, seed):
    """Synthetic mixtures: mixture + foreground target, CPU generator."""
    from _util import synthetic_mixture
    shape = (wl['batch'], wl['channels'], wl['seconds'] * FS)
    if wl['channels'] == 1:
        shape = (wl['batch'], wl['seconds'] * FS)
    mix, fg = synthetic_mixture(shape, seed)
    if fg.ndim == 3:
        fg = fg.mean(1)      # single-channel training target (ffnn.py:107, tfgridnet.py:137)
    return mix, fg


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
                            'cfg3': ['sisnr'],
                            'cfg4': ['stft', 'istft', 'sisnr'],
                            'cfg5': ['stft', 'istft', 'sisnr']}[name]

    # the stages, each one public-API call (plus the view the model takes between them)
    def stage_stft(self, mix):
        return self.stft(mix)

    def stage_features(self, spec):
        return self.front.features(spec, self.mean, self.std)

    def select(self, spec):
        """What goes back through the iSTFT: channel mean (FFNN), source 0 (TF-GridNet)."""
        if self.name == 'cfg1':
            return self.brv.ffnn.channel_mean(spec)      # FFNN._enhance's x.mean(1), one fused pass
        if self.name == 'cfg5':
            return spec[:, 0]
        return spec

    def stage_istft(self, spec):
        return self.stft.backward(spec)[..., :self.samples]

    def stage_sisnr(self, y, target):
        return self.brv.sisnr(y.unsqueeze(1), target.unsqueeze(1), self.lengths)

    def step(self, mix, target, marks=None):
        """mix/target: device tensors.  Returns the (batch,) loss tensor."""
        def mark(i):
            if marks is not None:
                marks[i].record()
        mark(0)
        if self.name == 'cfg3':          # the estimate comes from a learned decoder: criterion only
            loss = self.stage_sisnr(mix, target)
            mark(1)
            return loss
        spec = self.stage_stft(mix)
        mark(1)
        k = 2
        if self.name == 'cfg1':
            feats = self.stage_features(spec)   # noqa: F841
            mark(k)
            k += 1
        y = self.stage_istft(self.select(spec))
        mark(k)
        loss = self.stage_sisnr(y, target)
        mark(k + 1)
        return loss


def time_stages(pipe, sets, reps, only=None):
    """Device time of each stage in isolation: the stage's launches for every input
    set are captured in one CUDA graph and replayed `reps` times between two CUDA
    events on the launching stream (no host gaps; inputs rotate, so nothing is
    L2-resident from the previous launch of the same stage)."""
    from brever_b200 import graphs
    if pipe.name == 'cfg3':
        specs, sel, ys = [], [], [m for m, _ in sets]
    else:
        specs = [pipe.stage_stft(m) for m, _ in sets]
        sel = [pipe.select(sp) for sp in specs]
        ys = [pipe.stage_istft(sp) for sp in sel]
    calls = {
        'stft': lambda: [pipe.stage_stft(m) for m, _ in sets],
        'features': lambda: [pipe.stage_features(sp) for sp in specs],
        'istft': lambda: [pipe.stage_istft(sp) for sp in sel],
        'sisnr': lambda: [pipe.stage_sisnr(y, f) for y, (_, f) in zip(ys, sets)],
    }
    out = {}
    for name in pipe.stage_names:
        if only and name not in only:
            continue
        g = graphs.capture(calls[name])
        for _ in range(2):
            g()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / (reps * len(sets))
        del g
    return out


def measure_workload(name, wl, device, rank, world, steps, warmup, use_graph=True, with_e2e=True,
                     global_audio_s=None, seed_base=1000):
    """One workload through the public API on this rank's GPU: device-timed steps (max over
    ranks), per-stage times, the roofline block and the end-to-end (host buffers) figure.
    `global_audio_s`: audio-seconds all ranks process per step (default: world x this rank's)."""
    import torch.distributed as dist
    from brever_b200 import _lib, graphs
    peaks = load_peaks()
    pipe = Pipeline(name, wl, device)
    work, n_frames = stage_work(wl)
    audio_s = wl['batch'] * wl['seconds']            # this rank, per step
    if global_audio_s is None:
        global_audio_s = world * audio_s

    # rotating input sets so that a step never finds its inputs in L2
    set_bytes = sum(work[s]['bytes'] for s in pipe.stage_names)
    n_sets = max(2, min(8, int(2.2 * 126e6 / max(set_bytes, 1)) + 1))
    sets = []
    for i in range(n_sets):
        mix, fg = make_batch(wl, seed_base + rank * 16 + i)
        sets.append((mix.to(device), fg.to(device)))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total = torch.zeros((), device=device)

    def full_step(mix, target):
        loss = pipe.step(mix, target)
        pipe.brv.ffnn.accumulate_mean(total, loss)   # running metric (training.py:369-373), one launch
        return loss

    lib = _lib.lib()
    if use_graph:
        # one captured step per input set: a replay is ONE host call for the whole chain
        graphed = [graphs.capture(full_step, m, f) for m, f in sets]
        launches_per_step = graphed[0].launches

        def run(i):
            return graphed[i % n_sets]()
    else:
        c0 = lib.brv_launch_count()
        full_step(*sets[0])
        launches_per_step = int(lib.brv_launch_count() - c0)

        def run(i):
            return full_step(*sets[i % n_sets])

    for i in range(warmup):
        run(i)
    barrier()
    total.zero_()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    start.record()
    for i in range(steps):
        run(i)
    if world > 1:   # the one collective: final metric all-reduce (training.py:369-373)
        dist.all_reduce(total)
    stop.record()
    barrier()
    t_end = time.perf_counter()
    elapsed_ms = start.elapsed_time(stop)
    t = torch.tensor([elapsed_ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / steps
    value = global_audio_s / (ms_per_step * 1e-3)
    mean_loss = float(total) / (steps * world)

    # per-stage device time, each stage isolated (CUDA events around graph replays)
    stage_ms = time_stages(pipe, sets, reps=max(5, min(steps, 20)))
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
    try:    # DRAM traffic of the dominant kernel from the committed ncu capture of this round
        with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as f:
            tr = json.load(f)
        roofline['traffic'] = tr.get(name, {}).get(dominant)
        roofline['traffic_source'] = tr.get('_source')
    except (OSError, ValueError):
        pass
    roofline['stage_ms'] = {k: round(v, 4) for k, v in stage_ms.items()}
    roofline['stage_hbm_frac'] = {
        k: round(work[k]['bytes'] / (v * 1e-3) / 1e9 / peaks['hbm'], 4) for k, v in stage_ms.items()}
    roofline['stage_timing'] = 'each stage alone: CUDA events around CUDA-graph replays over the rotating input sets'
    # the whole chain against its own roofline: max(bytes / HBM, flops / (bf16 / 3))
    tot_bytes = sum(work[s]['bytes'] for s in pipe.stage_names)
    tot_flops = sum(work[s]['flops'] for s in pipe.stage_names)
    t_roof = max(tot_bytes / (peaks['hbm'] * 1e9), tot_flops / (peaks['bf16'] / 3 * 1e12))
    local_ms = ms_per_step                      # every rank runs the same amount of work
    roofline['chain'] = {'alg_bytes': tot_bytes, 'alg_flops': tot_flops,
                         'roofline_ms': round(t_roof * 1e3, 4),
                         'frac': round(t_roof * 1e3 / local_ms, 4)}
    res = {'ms_per_step': round(ms_per_step, 4), 'value': round(value, 1), 'unit': 'audio-s/s',
           'launches_per_step': int(launches_per_step), 'mean_loss_db': round(mean_loss, 4),
           'frames': n_frames, 'n_sets': n_sets, 'roofline': roofline,
           '_t': (t_begin, t_end), '_pipe': (pipe, sets, full_step)}
    if not with_e2e:
        return res

    # ---- end to end: pinned host buffers, H2D + D2H inside the timed region ----
    # mixture and target of a step travel as ONE pinned buffer / one H2D copy (they are
    # views of one device buffer on the other side): fewer DMA set-ups per step
    def pack(m, f):
        return torch.cat([m.reshape(-1), f.reshape(-1)]).cpu().pin_memory()

    cpus_before = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(device.index)     # before the pinned buffers are first touched
    host = [pack(m, f) for m, f in sets[:2]]
    n_mix = sets[0][0].numel()
    h2d = host[0].numel() * 4
    d2h = wl['batch'] * 4
    copy_stream = torch.cuda.Stream(device)
    dev_flat = [torch.empty(host[0].numel(), dtype=torch.float32, device=device) for _ in range(2)]
    dev_bufs = [(d[:n_mix].view(sets[0][0].shape), d[n_mix:].view(sets[0][1].shape)) for d in dev_flat]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    result = torch.empty(wl['batch'], dtype=torch.float32).pin_memory()
    os.sched_setaffinity(0, cpus_before)           # the CPU baseline legs use every core again
    if use_graph:
        e2e_steps = [graphs.capture(full_step, *dev_bufs[b]) for b in range(2)]
    else:
        e2e_steps = [lambda b=b: full_step(*dev_bufs[b]) for b in range(2)]

    def e2e_loop(n):
        main = torch.cuda.current_stream(device)
        for i in range(n + 1):
            if i < n:                      # stage the next step's inputs
                b = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[b])
                    dev_flat[b].copy_(host[i % 2], non_blocking=True)
                    ready[b].record(copy_stream)
            if i >= 1:                     # compute step i-1
                b = (i - 1) % 2
                main.wait_event(ready[b])
                loss = e2e_steps[b]()
                consumed[b].record(main)
                result.copy_(loss, non_blocking=True)   # D2H read of the step result
        main.synchronize()

    for ev in consumed:
        ev.record(torch.cuda.current_stream(device))
    e2e_loop(max(2, warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = global_audio_s * steps / float(t)
    # what the link alone allows: the same H2D copies with no compute behind them
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        dev_flat[i % 2].copy_(host[i % 2], non_blocking=True)
    torch.cuda.synchronize()
    h2d_only_s = (time.perf_counter() - t0) / steps
    rates = torch.tensor([h2d / h2d_only_s / 1e9], device=device)
    if world > 1:
        gathered = [torch.zeros_like(rates) for _ in range(world)]
        dist.all_gather(gathered, rates)
        rates = torch.cat(gathered)
    res['e2e'] = {'value': round(e2e_value, 1), 'unit': 'audio-s/s',
                  'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                  'h2d_only_ms_per_step': round(h2d_only_s * 1e3, 4),
                  'h2d_gb_per_s': round(h2d / h2d_only_s / 1e9, 1),
                  'h2d_gb_per_s_per_rank': [round(float(r), 1) for r in rates],
                  'numa': numa,
                  'note': 'one pinned host -> device copy of mixture+target per step, double-buffered on a copy stream; '
                          'loss read back every step; h2d_only_* = the same copies with no compute (the PCIe bound)'}
    return res


def torch_gpu_baseline(wl, device, steps):
    """The reference's own library calls for the cfg2 chain (torch.stft / torch.istft -> cuFFT,
    elementwise SI-SNR; brever/modules/stft.py:59-138, criterion.py:41-72 with one source) on
    CUDA tensors of the same B200: the existing-library bar."""
    eps = torch.finfo(torch.float32).eps
    N, H = wl['frame_length'], wl['hop']
    window = torch.hann_window(N, periodic=True, device=device)
    norm = window.pow(2).sum().sqrt()
    mix, fg = make_batch(wl, 1000)
    mix, fg = mix.to(device), fg.to(device)
    lengths = torch.full((wl['batch'],), mix.shape[-1], dtype=torch.int64, device=device)

    def step():
        spec = torch.stft(mix, n_fft=N, hop_length=H, win_length=N, window=window, center=True,
                          pad_mode='constant', normalized=False, onesided=True, return_complex=True) / norm
        y = torch.istft(spec * norm, n_fft=N, hop_length=H, win_length=N, window=window, center=True,
                        normalized=False, onesided=True, return_complex=False)[..., :mix.shape[-1]]
        mask = (torch.arange(mix.shape[-1], device=device)[None] < lengths[:, None]).float()
        a, b = y * mask, fg * mask
        a = (a - a.sum(-1, keepdim=True) / lengths[:, None]) * mask
        b = (b - b.sum(-1, keepdim=True) / lengths[:, None]) * mask
        proj = (a * b).sum(-1, keepdim=True) * b / b.pow(2).sum(-1, keepdim=True)
        noise = a - proj
        return -10 * torch.log10(proj.pow(2).sum(-1) / (noise.pow(2).sum(-1) + eps) + eps)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'value': round(wl['batch'] * wl['seconds'] / (ms * 1e-3), 1), 'unit': 'audio-s/s',
            'ms_per_step': round(ms, 4),
            'what': f'torch {torch.__version__} torch.stft / torch.istft (cuFFT) + elementwise SI-SNR on the same GPU, '
                    'eager launches, same inputs and shapes as the headline workload'}


# global batches of the fixed-batch (strong-scaling) configs of BASELINE.json
STRONG = {'cfg5': 1024, 'cfg3': 256}


def run_ours(args, rank, world, device):
    from brever_b200 import distributed
    wl = WORKLOADS[args.workload]
    sampler = ClockSampler(device.index)
    sampler.start()
    use_graph = not args.no_graph
    strong = args.scaling == 'strong'
    if strong:
        # fixed global batch, utterances sharded by rank (brever/batching.py:279-290)
        lo, hi = distributed.shard_bounds(wl['batch'], rank, world)
        wl = dict(wl, batch=hi - lo)
        global_audio = WORKLOADS[args.workload]['batch'] * wl['seconds']
    else:
        global_audio = None
    res = measure_workload(args.workload, wl, device, rank, world, args.steps, args.warmup,
                           use_graph=use_graph, global_audio_s=global_audio)
    t_begin, t_end = res.pop('_t')
    pipe, sets, full_step = res.pop('_pipe')
    clocks = sampler.stop(t_begin, t_end)
    n_sets = res['n_sets']
    base_wl = WORKLOADS[args.workload]
    out = {
        'metric': 'audio-seconds/sec (STFT->iSTFT->SI-SNR front-end)',
        'value': res['value'], 'unit': 'audio-s/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['ms_per_step'],
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"{args.workload}: {base_wl['desc']}",
                   'per_gpu_batch': wl['batch'], 'seconds': wl['seconds'], 'fs': FS,
                   'frame_length': wl['frame_length'], 'hop_length': wl['hop'],
                   'frames': res['frames'],
                   'parallelism': f'dp{world} (utterances sharded by rank' +
                                  (f', fixed global batch {base_wl["batch"]})' if strong else ', fixed per-GPU batch)'),
                   'l2': f'{n_sets} rotating input sets (> 2x L2) so steps never hit L2-resident inputs',
                   'launch': 'one CUDA-graph replay per step (brever_b200.graphs.capture of the public API calls)'
                             if use_graph else 'eager Python launches',
                   'stft_path': os.environ.get('BRV_FORCE_GENERIC', '0') == '1' and 'generic' or 'default'},
        'e2e': res['e2e'],
        'gpu_launches': int(res['launches_per_step'] * args.steps),
        'mean_loss_db': res['mean_loss_db'],
        'roofline': res['roofline'],
        'clocks': clocks,
    }
    if use_graph and args.eager_compare:
        # the same steps launched eagerly from Python, for the launch-overhead picture
        for i in range(3):
            full_step(*sets[i % n_sets])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            full_step(*sets[i % n_sets])
        e1.record()
        torch.cuda.synchronize()
        out['eager_ms_per_step'] = round(e0.elapsed_time(e1) / args.steps, 4)
    del pipe, sets, full_step
    torch.cuda.empty_cache()

    if world == 1 and not args.no_extras:
        # every BASELINE.json config in the same run (lighter settings), so that K2 (mel /
        # features, cfg1) and the 60 % target are measured by whoever runs this file
        sub_steps = max(10, min(args.steps, 20))
        table = {args.workload: {k: res[k] for k in ('ms_per_step', 'value', 'launches_per_step')} |
                 {'chain_frac': res['roofline']['chain']['frac'], 'dominant': res['roofline']['kernel'],
                  'dominant_frac': res['roofline']['frac'], 'stage_ms': res['roofline']['stage_ms'],
                  'e2e': res['e2e']['value']}}
        for name in sorted(WORKLOADS):
            if name in table:
                continue
            r = measure_workload(name, WORKLOADS[name], device, rank, world, sub_steps, 3,
                                 use_graph=use_graph)
            r.pop('_t'), r.pop('_pipe')
            table[name] = {'ms_per_step': r['ms_per_step'], 'value': r['value'],
                           'launches_per_step': r['launches_per_step'],
                           'chain_frac': r['roofline']['chain']['frac'],
                           'dominant': r['roofline']['kernel'], 'dominant_frac': r['roofline']['frac'],
                           'dominant_bound': r['roofline']['bound'],
                           'stage_ms': r['roofline']['stage_ms'], 'e2e': r['e2e']['value'],
                           'desc': WORKLOADS[name]['desc']}
            torch.cuda.empty_cache()
        out['workloads'] = table
        out['workloads_note'] = (f'all five BASELINE.json configs on this GPU in this run; {sub_steps} timed steps '
                                 'each besides the headline workload; chain_frac = step time against '
                                 'max(bytes / HBM, flops / (bf16 / 3)) of SURVEY 8d')
        if args.workload == 'cfg2':
            out['torch_gpu_baseline'] = torch_gpu_baseline(WORKLOADS['cfg2'], device, sub_steps)
    if world > 1 and not args.no_extras and not strong:
        # the fixed-batch configs of north_star, sharded with brever_b200.distributed
        block = {}
        for name, gbatch in STRONG.items():
            lo, hi = distributed.shard_bounds(gbatch, rank, world)
            swl = dict(WORKLOADS[name], batch=hi - lo)
            r = measure_workload(name, swl, device, rank, world, max(10, min(args.steps, 20)), 3,
                                 use_graph=use_graph, with_e2e=False,
                                 global_audio_s=gbatch * swl['seconds'])
            r.pop('_t'), r.pop('_pipe')
            block[name] = {'global_batch': gbatch, 'per_gpu_batch': hi - lo, 'ms_per_step': r['ms_per_step'],
                           'value': r['value'], 'chain_frac': r['roofline']['chain']['frac']}
            torch.cuda.empty_cache()
        out['strong_scaling'] = block
        out['strong_scaling_note'] = ('fixed global batch (cfg5: 1024 x 2ch x 4 s, cfg3: 256 x 4 s) sharded by rank with '
                                      'brever_b200.distributed.shard_bounds; value = global audio-s / max-over-ranks step time; '
                                      'compare with workloads.cfg5 / cfg3 of the 1-GPU line')
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_reference(args.workload, budget_s=15.0)
        one = cpu_reference(args.workload, budget_s=6.0, threads=1)
        out['cpu_baseline_1thread'] = {k: one[k] for k in ('value', 'unit', 'cores', 'sample')}
    if world > 1:
        out['vs_reference_note'] = ('the reference arm runs ONE batch slice on rank 0 host cores; at N GPUs the whole-job '
                                    'value covers N batches (weak scaling), so value / reference grows with N by construction')
    return out


# --------------------------------------------------------------------------- #
# reference arm: the reference's CPU implementation on this host's cores      #
# --------------------------------------------------------------------------- #
def cpu_reference(workload, budget_s=15.0, steps=None, warmup=1, threads=None):
    """Times oracle/torch_port.py (the reference's exact torch calls, float32 CPU,
    all host threads) on a bounded slice of the workload."""
    from oracle import tf_oracle as O
    from oracle import torch_port as P
    wl = WORKLOADS[workload]
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    win = torch.from_numpy(O.get_window('hann', wl['frame_length']))
    kw = dict(frame_length=wl['frame_length'], hop_length=wl['hop'], **wl['kw'])
    filters = torch.from_numpy(O.mel_filterbank()[0]) if workload == 'cfg1' else None
    S = wl['seconds'] * FS

    def run(mix, fg):
        with torch.no_grad():
            if workload == 'cfg3':
                lengths = torch.full((mix.shape[0],), S)
                return P.sisnr(mix.unsqueeze(1), fg.unsqueeze(1), lengths)
            spec = P.stft(mix, win, **kw)
            if workload == 'cfg1':
                feats = P.stack(P.logfbe(spec, filters), 5)
                feats = P.static_normalize(feats, 0.0, 1.0)  # noqa: F841
                spec = spec.mean(1)
            elif workload == 'cfg5':
                spec = spec[:, 0]
            y = P.istft(spec, win, **kw)[..., :S]
            lengths = torch.full((mix.shape[0],), S)
            return P.sisnr(y.unsqueeze(1), fg.unsqueeze(1), lengths)

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
    _, n_frames = stage_work(wl)
    return {
        'impl': 'reference',
        'metric': 'audio-seconds/sec (STFT->iSTFT->SI-SNR front-end)',
        'value': base['value'], 'unit': 'audio-s/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': base['ms_per_pass'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"{args.workload}: {wl['desc']}", 'per_gpu_batch': wl['batch'],
                   'seconds': wl['seconds'], 'fs': FS, 'frame_length': wl['frame_length'],
                   'hop_length': wl['hop'], 'frames': n_frames,
                   'parallelism': 'host CPU cores of rank 0 (no GPU)', 'l2': 'n/a (CPU arm)',
                   'launch': 'eager torch CPU calls', 'stft_path': 'reference (torch.stft / torch.istft, MKL FFT)',
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
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: every rank runs the full per-GPU batch; strong: the fixed global batch of the '
                         'workload is sharded over the ranks (brever_b200.distributed.shard_bounds)')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the per-config table (N = 1), the torch-on-GPU baseline and the strong-scaling block (N > 1)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch eagerly from Python instead of replaying CUDA graphs')
    ap.add_argument('--no-eager-compare', dest='eager_compare', action='store_false',
                    help='skip the extra eagerly-launched pass reported as eager_ms_per_step')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    # exactly ONE line on stdout: libraries that print there (NCCL's version banner does) are
    # sent to stderr, the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + '\n').encode())

    if args.impl == 'reference':
        out = run_reference(args, rank, world)
        if out is not None:
            emit(out)
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
        emit(out)
    return 0


if __name__ == '__main__':
    sys.exit(main())
