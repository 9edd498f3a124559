#!/usr/bin/env python
"""bench.py -- voice-samples/sec of the block-render hot path on N B200s of one node.

Workload (BASELINE.json configs[1], "C2"): sine -> biquad (Butterworth) low-pass -> gain chain,
4,096 independent voices x 10 s at 48 kHz per GPU, float32 output (frames, voices) materialised in
HBM.  One *step* = one full render of that block.  N>1 shards voices across ranks (each rank owns
its own 4,096-voice bank; no data-path collective) => weak scaling.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU (numpy/scipy) path, host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RATE = 48000
METRIC = 'voice-samples/sec'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--voices', type=int, default=4096)
    ap.add_argument('--seconds', type=float, default=10.0)
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scan-variant', type=int, default=None)
    ap.add_argument('--plan-opt', action='append', default=[], help='key=value passed to sigb_plan_set_option (A/B testing)')
    return ap.parse_args()


def config(args, n):
    return {'workload': 'C2: sine -> biquad lowpass -> gain, %d voices x %g s @ 48 kHz per GPU, fp32 (frames, voices) block in HBM'
                        % (args.voices, args.seconds),
            'voices_per_gpu': args.voices, 'frames': int(args.seconds * RATE), 'rate': RATE,
            'sharding': 'voices across %d rank(s), no collective' % n,
            'l2': 'output block (%.2f GB) >> 126 MB L2, rewritten every step; no flush needed'
                  % (args.voices * args.seconds * RATE * 4 / 1e9)}


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's numpy/scipy render (never on the product path)
# ------------------------------------------------------------------------------------------------
def _cpu_chunk(job):
    from oracle import np_oracle
    hertz, phase, cutoff, g, frames = job
    out = np_oracle.render_voice_chain(0, frames, RATE, hertz, phase, cutoff, g)
    return float(out[-1].sum())


def cpu_render(voices, frames, workers, seed=2):
    """Reference render (oracle port) of `voices` voices x `frames` frames fanned over `workers`
    processes by channel chunk; returns seconds."""
    from oracle import cases
    hertz, phase, cutoff, g = cases.voice_params(seed, voices)
    per = max(1, min(64, voices // max(1, workers)))
    jobs = [(hertz[i:i + per], phase[i:i + per], cutoff[i:i + per], g[i:i + per], frames)
            for i in range(0, voices, per)]
    t0 = time.perf_counter()
    if workers <= 1:
        for j in jobs:
            _cpu_chunk(j)
    else:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(workers) as pool:
            pool.map(_cpu_chunk, jobs, chunksize=1)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = cores
    frames = RATE            # bounded sample: 1 s of every sampled voice
    voices = max(workers * 16, 64)
    # pool start-up is part of neither arm's steady state: time the pool-resident render only
    from oracle import cases
    import multiprocessing as mp
    hertz, phase, cutoff, g = cases.voice_params(2, voices)
    per = 16
    jobs = [(hertz[i:i + per], phase[i:i + per], cutoff[i:i + per], g[i:i + per], frames)
            for i in range(0, voices, per)]
    times = []
    with mp.get_context('fork').Pool(workers) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_chunk, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = voices * frames * len(times) / total
    sample = '%d of %d voices x 1 s per step (single-request render per %d-voice chunk), %d processes' % (
        voices, args.voices, per, workers)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voice-samples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config(args, args.gpus),
            'cpu_baseline': {'value': value, 'unit': 'voice-samples/s', 'cores': workers, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'voice-samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': 'numpy/scipy oracle port of the reference render (the Python reference cannot travel to the GPU box)'}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:   # noqa: BLE001
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.002)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples), 'source': 'nvml' if self.ok else 'unavailable'}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from oracle import cases            # parameter distributions only (shared with the tests)
    from signals_b200 import engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the block render has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    v, frames = args.voices, int(args.seconds * RATE)
    ns = cases.b200_namespace()
    hertz, phase, cutoff, g = cases.voice_params(2 + rank, v)

    def build_graph():
        return cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])

    eng = engine.Engine(device=torch.device('cuda', local))
    compiled = eng.compile(build_graph(), v, RATE, frames)
    if args.scan_variant is not None:
        compiled.set_option('scan_variant', args.scan_variant)
    for kv in args.plan_opt:
        k, val = kv.split('=')
        compiled.set_option(k, int(val))
    out = torch.empty((frames, v), dtype=torch.float32, device='cuda')

    for _ in range(max(args.warmup, 3)):
        compiled.render_device(0, frames, out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = compiled.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kernel_ms = []
    ev[0].record()
    for i in range(args.steps):
        compiled.render_device(0, frames, out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    kernel_ms.append(compiled.last_kernel_ms())
    clocks = sampler.finish()
    barrier()
    launches = compiled.launch_count - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    units = float(world) * v * frames * args.steps
    value = units / (total_ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers: compile (host->device tables) +
    #      render_host (kernels + pipelined device->host copies), every step
    host_out = torch.empty((frames, v), dtype=torch.float32, pin_memory=True)
    e2e_times = []
    param_bytes = compiled.describe()['param_bytes']
    e2e_launches = 0
    for i in range(args.e2e_steps + 1):
        barrier()
        t0 = time.perf_counter()
        c2 = eng.compile(build_graph(), v, RATE, frames)
        if args.scan_variant is not None:
            c2.set_option('scan_variant', args.scan_variant)
        c2.render_host(0, frames, host_out)
        dt = time.perf_counter() - t0
        e2e_launches = c2.launch_count
        c2.close()
        if i > 0:
            e2e_times.append(dt)
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(world) * v * frames * len(e2e_times) / float(te.item()) if e2e_times else None
    checksum = float(host_out[-1].double().sum())

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        per_launch_bytes = 4.0 * v * frames                      # algorithmic: one fp32 store per voice-sample
        avg_ms = float(np.mean(step_ms))
        achieved = per_launch_bytes / (avg_ms * 1e-3) / 1e9
        line = {'metric': METRIC, 'value': value, 'unit': 'voice-samples/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (fp64 phase / Q0.64 phase accumulator, fp64 scan carries)',
                'data': 'synthetic', 'config': config(args, world),
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                             'traffic': None, 'kernel': 'k_chain_scan (+ k_chain_seq tail rows)',
                             'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else '6650 GB/s (of fallback)',
                             'algorithmic_bytes_per_voice_sample': 4, 'last_render_ms_in_library': kernel_ms[-1]},
                'e2e': {'value': e2e_value, 'unit': 'voice-samples/s', 'h2d_bytes_per_step': int(param_bytes),
                        'd2h_bytes_per_step': int(4 * v * frames), 'steps': len(e2e_times),
                        'what': 'Engine.compile(graph) + CompiledPlan.render_host(pinned fp32 block)'},
                'gpu_launches': int(launches), 'gpu_launches_e2e_per_step': int(e2e_launches),
                'clocks': clocks, 'step_ms_min': float(np.min(step_ms)), 'step_ms_max': float(np.max(step_ms)),
                'checksum_last_frame': checksum}
        if world == 1 and not args.no_cpu_baseline:
            sample_v, sample_f = 192, RATE * 10
            secs = cpu_render(sample_v, sample_f, workers=1)
            line['cpu_baseline'] = {'value': sample_v * sample_f / secs, 'unit': 'voice-samples/s', 'cores': 1, 'kind': 'port',
                                    'sample': '%d of %d voices x 10 s, single request, 1 process (the reference is single-threaded); %.1f s of CPU'
                                              % (sample_v, v, secs),
                                    'host_cores': os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
