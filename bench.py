#!/usr/bin/env python
"""bench.py -- voice-samples/sec of the block-render hot path on N B200s of one node.

Headline workload (BASELINE.json configs[1], "C2"): sine -> biquad (Butterworth) low-pass -> gain,
4,096 independent voices x 10 s at 48 kHz per GPU, float32 output (frames, voices) materialised in
HBM.  One *step* = one full render of that block.  N>1 shards voices across ranks (each rank owns its
own 4,096-voice bank; no data-path collective) => weak scaling.

Other BASELINE configs (extra lines, same JSON contract): --config c3 (additive bank, 65,536 sine
partials -> 64 channels, fused oscillator + mix reduction), --config c4 (8-biquad cascade, 16,384
channels x 60 s streamed in 1 s slabs with carried state), --config c5 (1M randomised instances
sharded by voice over the ranks, NCCL reduce of the stereo mix-down; strong scaling).

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
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--voices', type=int, default=None, help='voices / partials / channels / instances (config default if omitted)')
    ap.add_argument('--seconds', type=float, default=None)
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--slab-seconds', type=float, default=10.0, help='c4: seconds of audio per streamed slab')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scan-variant', type=int, default=None)
    ap.add_argument('--plan-opt', action='append', default=[], help='key=value passed to sigb_plan_set_option (A/B testing)')
    ap.add_argument('--default-opt', action='append', default=[], help='key=value passed to sigb_set_default_option')
    args = ap.parse_args()
    defaults = {'c2': (4096, 10.0, 20, 3), 'c3': (65536, 10.0, 5, 1), 'c4': (16384, 60.0, 2, 1), 'c5': (1 << 20, 10.0, 3, 1)}
    v, s, k, e = defaults[args.config]
    args.voices = args.voices if args.voices is not None else v
    args.seconds = args.seconds if args.seconds is not None else s
    args.steps = args.steps if args.steps is not None else k
    args.e2e_steps = args.e2e_steps if args.e2e_steps is not None else e
    return args


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE config: graph builder, unit accounting, roofline model, CPU sample."""
    scaling = 'weak'
    out_channels = None
    slab_frames = None           # render in slabs of this many frames (streamed configs)

    def __init__(self, args, rank, world):
        self.args, self.rank, self.world = args, rank, world
        self.frames = int(args.seconds * RATE)

    def units_per_step(self):     # whole job, all ranks
        raise NotImplementedError


class C2(Workload):
    name = 'C2: sine -> biquad lowpass -> gain, %d voices x %g s @ 48 kHz per GPU, fp32 (frames, voices) block in HBM'
    kernel = 'k_chain_scan3<sine, 1 section, 64-channel tiles, 3x9 workers, f32 carry chain>'
    traffic_profile = 'r01_k_chain_scan3_full.txt'
    bound = 'hbm'
    bytes_per_unit = 4.0

    def __init__(self, args, rank, world):
        super().__init__(args, rank, world)
        from signals_b200 import workloads as cases
        self.v = args.voices
        self.out_channels = self.v
        self.params = cases.voice_params(2 + rank, self.v)

    def build(self, ns):
        from signals_b200 import workloads as cases
        hertz, phase, cutoff, g = self.params
        return cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])

    def describe(self):
        return {'workload': self.name % (self.v, self.args.seconds), 'voices_per_gpu': self.v, 'frames': self.frames, 'rate': RATE,
                'sharding': 'voices across %d rank(s), no collective' % self.world,
                'l2': 'output block (%.2f GB) >> 126 MB L2, rewritten every step; no flush needed' % (self.v * self.frames * 4 / 1e9)}

    def units_per_step(self):
        return float(self.world) * self.v * self.frames

    def launch_units(self):
        return float(self.v) * self.frames

    def cpu_sample(self, workers):
        from signals_b200 import workloads as cases
        sv = 192 if workers == 1 else max(workers * 16, 64)
        sf = RATE * 10 if workers == 1 else RATE
        hertz, phase, cutoff, g = cases.voice_params(2, sv)
        per = 64 if workers == 1 else 16
        jobs = [('chain', (hertz[i:i + per], phase[i:i + per], cutoff[i:i + per], g[i:i + per], sf)) for i in range(0, sv, per)]
        return jobs, sv * sf, '%d of %d voices x %g s, single request per %d-voice chunk' % (sv, self.v, sf / RATE, per)


class C3(Workload):
    name = 'C3: additive bank, %d sine partials -> %d channels (fused oscillator + mix reduction), %g s @ 48 kHz per GPU'
    kernel = 'k_bank'
    bound = 'sfu'
    bytes_per_unit = 4.0 / 1024

    def __init__(self, args, rank, world):
        super().__init__(args, rank, world)
        from signals_b200 import workloads as cases
        self.p = args.voices
        self.groups = max(1, self.p // 1024)
        self.out_channels = self.groups
        self.params = cases.bank_params(3 + rank, self.p, self.p // self.groups)

    def build(self, ns):
        from signals_b200 import workloads as cases
        from signals_b200.chain import ext
        return cases.build_bank(ns, ext, *self.params, self.groups)

    def describe(self):
        return {'workload': self.name % (self.p, self.groups, self.args.seconds), 'partials_per_gpu': self.p, 'frames': self.frames,
                'rate': RATE, 'sharding': 'banks across %d rank(s), no collective' % self.world,
                'l2': 'compute-bound (MUFU): parameters 1.5 MB, output %.0f MB rewritten every step' % (self.groups * self.frames * 4 / 1e6)}

    def units_per_step(self):
        return float(self.world) * self.p * self.frames

    def launch_units(self):
        return float(self.p) * self.frames

    def cpu_sample(self, workers):
        from signals_b200 import workloads as cases
        sp, sf = 1024 * max(1, min(workers, 8)), RATE // 2
        hertz, phase, amp = cases.bank_params(3, sp, 1024)
        jobs = [('bank', (hertz[i:i + 1024], phase[i:i + 1024], amp[i:i + 1024], sf)) for i in range(0, sp, 1024)]
        return jobs, sp * sf, '%d of %d partials x %g s (one 1024-partial group per job)' % (sp, self.p, sf / RATE)


class C4(Workload):
    name = 'C4: 8-biquad low-pass cascade on %d channels x %g s @ 48 kHz per GPU, streamed in %g s slabs with carried state'
    kernel = 'k_cascade_reg (two channels per thread, all 8 sections in registers, equal time pieces per warp slot)'
    traffic_profile = 'r01_k_cascade_reg_full.txt'
    bound = 'hbm'
    bytes_per_unit = 8.0

    def __init__(self, args, rank, world):
        super().__init__(args, rank, world)
        self.ch = args.voices
        self.out_channels = self.ch
        self.slab_frames = min(self.frames, int(args.slab_seconds * RATE))
        rng = np.random.default_rng(4 + rank)
        self.cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, self.ch)))
        self.seed = 4 + rank

    def build(self, ns):
        import torch
        from signals_b200 import workloads as cases
        from signals_b200.chain import ext
        g = torch.Generator(device='cuda')
        g.manual_seed(self.seed)
        self.noise = torch.rand((self.slab_frames, self.ch), generator=g, device='cuda', dtype=torch.float32) * 2 - 1
        self.buffer = ext.Buffer(self.noise)
        node = self.buffer
        for s in range(8):
            node = cases.lowpass(ns, node, [self.cut[s]])
        return node

    def describe(self):
        return {'workload': self.name % (self.ch, self.args.seconds, self.slab_frames / RATE), 'channels_per_gpu': self.ch, 'frames': self.frames, 'rate': RATE,
                'slab_frames': self.slab_frames, 'sharding': 'channels across %d rank(s), no collective' % self.world,
                'l2': 'slab in + out = %.2f GB >> 126 MB L2; the same slab of U(-1,1) noise is re-bound at each slab position' 
                      % (2 * self.ch * self.slab_frames * 4 / 1e9)}

    def units_per_step(self):
        return float(self.world) * self.ch * self.frames

    def launch_units(self):
        return float(self.ch) * self.slab_frames

    def cpu_sample(self, workers):
        sc, sf = 16 * max(1, workers), RATE
        rng = np.random.default_rng(4)
        cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, sc)))
        jobs = [('cascade', (cut[:, i:i + 16], sf, 4 + i)) for i in range(0, sc, 16)]
        return jobs, sc * sf, '%d of %d channels x %g s through 8 chained LowPass nodes' % (sc, self.ch, sf / RATE)


class C5(Workload):
    name = 'C5: %d randomised osc/filter/gain/pan instances -> stereo mix, %g s @ 48 kHz, sharded by voice (instance i on rank i %% N)'
    kernel = 'k_voices (+ k_voices_finish)'
    bound = 'sfu'
    bytes_per_unit = 0.0
    scaling = 'strong'
    out_channels = 2

    def __init__(self, args, rank, world):
        super().__init__(args, rank, world)
        from signals_b200 import workloads as cases
        self.n = args.voices
        self.prm = cases.instance_params(5, self.n, rank, world)

    def build(self, ns):
        from signals_b200 import workloads as cases
        from signals_b200.chain import ext
        return cases.build_instances(ns, ext, self.prm)

    def describe(self):
        return {'workload': self.name % (self.n, self.args.seconds), 'instances_total': self.n, 'frames': self.frames, 'rate': RATE,
                'sharding': 'instance i on rank i %% %d; one reduce (NCCL) of the (frames, 2) mix per step' % self.world,
                'l2': 'compute-bound; parameter tables %.0f MB per GPU stream from L2/HBM once per step' % (self.n / self.world * 52 / 1e6)}

    def units_per_step(self):
        return float(self.n) * self.frames

    def launch_units(self):
        return float(len(self.prm['hertz'])) * self.frames

    def cpu_sample(self, workers):
        from signals_b200 import workloads as cases
        sn, sf = 64 * max(1, workers), RATE
        prm = cases.instance_params(5, sn)
        jobs = []
        for i in range(0, sn, 64):
            jobs.append(('instances', ({k: (v[i:i + 64] if isinstance(v, np.ndarray) else v) for k, v in prm.items()}, sf)))
        return jobs, sn * sf, '%d of %d instances x %g s' % (sn, self.n, sf / RATE)


WORKLOADS = {'c2': C2, 'c3': C3, 'c4': C4, 'c5': C5}


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's numpy/scipy render (never on the product path)
# ------------------------------------------------------------------------------------------------
def _cpu_job(job):
    from oracle import np_oracle
    kind, a = job
    if kind == 'chain':
        hertz, phase, cutoff, g, frames = a
        out = np_oracle.render_voice_chain(0, frames, RATE, hertz, phase, cutoff, g)
    elif kind == 'bank':
        hertz, phase, amp, frames = a
        out = np_oracle.render_bank(0, frames, RATE, hertz, phase, amp, 1)
    elif kind == 'cascade':
        cut, frames, seed = a
        x = np.random.default_rng(seed).uniform(-1, 1, (frames, cut.shape[1]))
        out, _ = np_oracle.render_cascade(x, cut, RATE)
    else:
        prm, frames = a
        out = np_oracle.render_instances(prm, 0, frames, RATE)
    return float(out[-1].sum())


def _cpu_init():
    from oracle import np_oracle   # noqa: F401  (numpy/scipy import cost is not part of the render)


def cpu_time(jobs, workers, pool=None):
    t0 = time.perf_counter()
    if workers <= 1:
        for j in jobs:
            _cpu_job(j)
    else:
        pool.map(_cpu_job, jobs, chunksize=1)
    return time.perf_counter() - t0


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: the Python reference cannot travel
    to the GPU box), all host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    workers = os.cpu_count() or 1
    wl = WORKLOADS[args.config](args, 0, max(1, args.gpus))
    jobs, units, sample = wl.cpu_sample(workers)
    times = []
    _cpu_init()
    with mp.get_context('fork').Pool(workers, initializer=_cpu_init) as pool:     # pool start-up is not part of the steady state
        for step in range(args.warmup + args.steps):
            dt = cpu_time(jobs, workers, pool)
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = units * len(times) / total
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voice-samples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': wl.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': wl.describe(),
            'cpu_baseline': {'value': value, 'unit': 'voice-samples/s', 'cores': workers, 'kind': 'port',
                             'sample': sample + ', %d processes' % workers},
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
def ncu_traffic(name):
    """dram__bytes_{read,write}.sum of a committed ncu summary under profiles/ (tools/ncu_summary.py format)."""
    if not name:
        return None
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    got = {}
    try:
        with open(os.path.join(ROOT, 'profiles', name)) as f:
            for ln in f:
                parts = ln.split()
                if len(parts) == 3 and parts[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and parts[2] in scale:
                    got[parts[0].split('_')[-1].split('.')[0]] = float(parts[1]) * scale[parts[2]]
    except OSError:
        return None
    return got if len(got) == 2 else None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from signals_b200 import workloads as cases     # parameter distributions and graph builders (shared with the tests)
    from signals_b200 import _lib, engine, shard

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

    for kv in args.default_opt:
        k, val = kv.split('=')
        assert _lib.lib().sigb_set_default_option(k.encode(), int(val)) == 0, kv

    wl = WORKLOADS[args.config](args, rank, world)
    frames = wl.frames
    ns = cases.b200_namespace()
    eng = engine.Engine(device=torch.device('cuda', local))
    graph = wl.build(ns)
    compiled = eng.compile(graph, wl.out_channels, RATE, frames)

    def configure(c):
        if args.scan_variant is not None:
            c.set_option('scan_variant', args.scan_variant)
        for kv in args.plan_opt:
            k, val = kv.split('=')
            c.set_option(k, int(val))

    configure(compiled)
    slab = wl.slab_frames or frames
    out = torch.empty((slab, wl.out_channels), dtype=torch.float32, device='cuda')
    reduce_mix = args.config == 'c5'

    def step(c):
        """One pass of the hot path over the whole workload, all on torch's current stream."""
        for r in range(0, frames, slab):
            if wl.slab_frames:
                c.bind_window(wl.buffer, wl.noise, r)
            c.render_device(r, min(slab, frames - r), out)
        if reduce_mix:
            shard.reduce_mix(out, dst=0)

    for _ in range(max(args.warmup, 3)):
        step(compiled)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = compiled.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step(compiled)
        ev[i + 1].record()
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    last_ms = compiled.last_kernel_ms()          # device time of the last sigb_render call (library's own events)
    clocks = sampler.finish()
    barrier()
    launches = compiled.launch_count - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = wl.units_per_step() * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers: compile (host->device tables) +
    #      render_host (kernels + pipelined device->host copies), every step
    # streamed workloads: the host-buffer leg moves 2.5 s slabs (7.9 GB pinned each way for C4) -- it is bound by
    # the PCIe copies, and the pinned staging stays small next to the device-resident slabs of the timed leg
    eslab = min(slab, int(2.5 * RATE)) if wl.slab_frames else slab
    host_out = torch.empty((eslab, wl.out_channels), dtype=torch.float32, pin_memory=True)
    host_in = wl.noise[:eslab].cpu().pin_memory() if wl.slab_frames else None
    e2e_times = []
    param_bytes = compiled.describe()['param_bytes']
    h2d = int(param_bytes)
    e2e_launches = 0
    for i in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        barrier()
        t0 = time.perf_counter()
        c2 = eng.compile(graph, wl.out_channels, RATE, frames)
        configure(c2)
        for r in range(0, frames, eslab):
            if wl.slab_frames:
                dev_in = host_in.to('cuda', non_blocking=True)        # this slab's input: pinned host -> HBM
                c2.bind_window(wl.buffer, dev_in, r)
            c2.render_host(r, min(eslab, frames - r), host_out)
        if reduce_mix and world > 1:
            mix = host_out.to('cuda', non_blocking=True)
            shard.reduce_mix(mix, dst=0)
            host_out.copy_(mix)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_launches = c2.launch_count
        c2.close()
        if i > 0:
            e2e_times.append(dt)
    if wl.slab_frames:
        h2d += int(host_in.numel() * 4 * ((frames + eslab - 1) // eslab))
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = wl.units_per_step() * len(e2e_times) / float(te.item()) if e2e_times else None
    checksum = float(host_out[-1].double().sum()) if e2e_times else float(out[-1].double().sum())

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except OSError:
            pass
        n_renders = (frames + slab - 1) // slab
        avg_step_ms = float(np.mean(step_ms))
        launch_ms = avg_step_ms / n_renders                       # one dominant-kernel launch per render call
        if wl.bound == 'hbm':
            peak = float(peaks.get('hbm_gbs', 6650.0))
            achieved = wl.bytes_per_unit * wl.launch_units() / (launch_ms * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else '6650 GB/s (of fallback)',
                    'algorithmic_bytes_per_voice_sample': wl.bytes_per_unit,
                    'algorithmic_bytes_per_launch': wl.bytes_per_unit * wl.launch_units()}
            # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this very
            # workload (never measured under the profiler here): only quoted when the launch has the captured size
            traffic = ncu_traffic(getattr(wl, 'traffic_profile', None))
            if traffic and abs(traffic['write'] / (4.0 * wl.launch_units()) - 1.0) < 0.02:
                roof['traffic'] = traffic['read'] + traffic['write']
                roof['traffic_unit'] = 'bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)'
                roof['traffic_source'] = 'profiles/' + wl.traffic_profile
        else:
            # transcendental-bound kernels: one MUFU.SIN per unit on the 16-lane/clk/SM special-function pipe
            sm_mhz = clocks.get('sm_mhz') or peaks.get('sm_max_mhz', 1965.0)
            peak = 148 * 16 * sm_mhz * 1e6 / 1e9
            achieved = wl.launch_units() / (launch_ms * 1e-3) / 1e9
            roof = {'bound': 'sfu', 'achieved': achieved, 'peak': peak, 'unit': 'Gsample/s', 'frac': achieved / peak, 'traffic': None,
                    'peak_source': '148 SMs x 16 MUFU lanes/clk x measured SM clock (derived; no MEASURED_PEAKS entry for the SFU pipe)',
                    'algorithmic_bytes_per_voice_sample': wl.bytes_per_unit}
        roof['kernel'] = wl.kernel
        roof['launch_ms'] = launch_ms
        roof['last_render_ms_in_library'] = last_ms
        line = {'metric': METRIC, 'value': value, 'unit': 'voice-samples/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True,
                'scaling': wl.scaling, 'vs_baseline': None, 'dtype': 'f32 (fp64 phase / Q0.64 phase accumulator, fp64 scan carries)',
                'data': 'synthetic', 'config': wl.describe(), 'roofline': roof,
                'e2e': {'value': e2e_value, 'unit': 'voice-samples/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': int(4 * wl.out_channels * frames), 'steps': len(e2e_times),
                        'what': 'Engine.compile(graph) + CompiledPlan.render_host(pinned fp32 block)'},
                'gpu_launches': int(launches), 'gpu_launches_e2e_per_step': int(e2e_launches),
                'clocks': clocks, 'step_ms_min': float(np.min(step_ms)), 'step_ms_max': float(np.max(step_ms)),
                'checksum_last_frame': checksum}
        if world == 1 and not args.no_cpu_baseline:
            jobs, units, sample = wl.cpu_sample(1)
            _cpu_init()
            secs = cpu_time(jobs, 1)
            line['cpu_baseline'] = {'value': units / secs, 'unit': 'voice-samples/s', 'cores': 1, 'kind': 'port',
                                    'sample': sample + ', 1 process (the reference is single-threaded); %.1f s of CPU' % secs,
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
