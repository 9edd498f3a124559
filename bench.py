#!/usr/bin/env python
"""bench.py -- voice-samples/sec of the block-render hot path on N B200s of one node.

Headline workload (BASELINE.json configs[1], "C2"): sine -> biquad (Butterworth) low-pass -> gain,
4,096 independent voices x 10 s at 48 kHz per GPU, float32 output (frames, voices) materialised in
HBM.  One *step* = one full render of that block.  N>1 shards voices across ranks (each rank owns its
own 4,096-voice bank; no data-path collective) => weak scaling.

The default run also measures the other BASELINE configs and appends them to the line as `extra`:
  extra.c5  1M randomised instances sharded by voice over the N ranks (instance i on rank i % N), ONE
            NCCL reduce of the (frames, 2) mix per step -- strong scaling; the reduce is also timed alone, and
            the N-rank mix is compared with rank 0's own render of the whole bank (`parity_n_vs_1`).
  extra.c3  (N=1) additive bank, 65,536 sine partials -> 64 channels (fused oscillator + mix reduction)
  extra.c4  (N=1) 8-biquad cascade, 16,384 channels x 60 s streamed in 10 s slabs with carried state
  extra.c2m (N=1) C2 with every cutoff driven by an LFO emitter: same time-parallel kernel, no cliff (`vs_unmodulated_c2`)
  extra.c1  (N=1) the audio callback: p50/p99 latency per block at 128/384/512/1024 frames for
            Sine<-Fixed -> Gain (scripts/example_sine.py as a graph) and for lowpass_test.sigs, through
            SinkDevice.render_block, with the reference's blockwise CPU render timed beside it.
Each of them can also be the headline: --config c3 | c4 | c5 (same JSON contract), --config c1 prints the latency
table alone.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU (numpy/scipy) path, host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import gc
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
# (voices, seconds, steps, e2e steps) when a config is the headline / when it rides along in `extra`
DEFAULTS = {'c2': (4096, 10.0, 20, 10), 'c2m': (4096, 10.0, 10, 0), 'c3': (65536, 10.0, 5, 3), 'c4': (16384, 60.0, 2, 1), 'c5': (1 << 20, 10.0, 3, 2)}
EXTRA_STEPS = {'c3': 5, 'c4': 2, 'c5': 3, 'c2m': 20}
EXTRA_WARMUP = {'c2m': 10}      # millisecond steps right after a CPU-only phase: let the clocks come back up first
FMA_PROBE_CEILING = 100.0       # FP32 lane-ops / clk / SM the delta-form section arithmetic reaches register-only (profiles/r02_fma_probe.txt; 96 for the state-variable form)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c1', 'c2', 'c2m', 'c3', 'c4', 'c5'])
    ap.add_argument('--voices', type=int, default=None, help='voices / partials / channels / instances (config default if omitted)')
    ap.add_argument('--seconds', type=float, default=None)
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--slab-seconds', type=float, default=10.0, help='c4: seconds of audio per streamed slab')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='headline config only')
    ap.add_argument('--extra', default='c2m,c5,c3,c4,c1', help='configs appended to the default (c2) line')
    ap.add_argument('--scan-variant', type=int, default=None)
    ap.add_argument('--plan-opt', action='append', default=[], help='key=value passed to sigb_plan_set_option (A/B testing)')
    ap.add_argument('--default-opt', action='append', default=[], help='key=value passed to sigb_set_default_option')
    args = ap.parse_args()
    v, s, k, e = DEFAULTS.get(args.config, DEFAULTS['c2'])
    args.voices_given = args.voices is not None
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

    def __init__(self, voices, seconds, rank, world, slab_seconds=10.0):
        self.voices, self.seconds, self.rank, self.world = voices, seconds, rank, world
        self.slab_seconds = slab_seconds
        self.frames = int(seconds * RATE)

    def units_per_step(self):     # whole job, all ranks
        raise NotImplementedError


class C2(Workload):
    name = 'C2: sine -> biquad lowpass -> gain, %d voices x %g s @ 48 kHz per GPU, fp32 (frames, voices) block in HBM'
    kernel = 'k_chain_scan3<sine, 1 section, 64-channel tiles, 3x9 workers, f32 carry chain>'
    traffic_profile = 'r02_k_chain_scan3_full.txt'
    bound = 'hbm'
    bytes_per_unit = 4.0

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        from signals_b200 import workloads as cases
        self.v = self.voices
        self.out_channels = self.v
        self.params = cases.voice_params(2 + self.rank, self.v)

    def build(self, ns):
        from signals_b200 import workloads as cases
        hertz, phase, cutoff, g = self.params
        return cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])

    def describe(self):
        return {'workload': self.name % (self.v, self.seconds), 'voices_per_gpu': self.v, 'frames': self.frames, 'rate': RATE,
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

    def cpu_sample_blockwise(self, workers):
        """The reference's PRODUCT mode: 512-frame callbacks, every filter re-designed per channel per block and run over
        100-frame context (fx.py:93-105)."""
        from signals_b200 import workloads as cases
        per = 8
        sv, sf = per * max(1, workers), RATE
        hertz, phase, cutoff, g = cases.voice_params(2, sv)
        jobs = [('blockwise', (hertz[i:i + per], phase[i:i + per], cutoff[i:i + per], g[i:i + per], sf, 512)) for i in range(0, sv, per)]
        return jobs, sv * sf, '%d of %d voices x %g s in 512-frame requests (per-block butter() per channel + 100-frame context)' % (sv, self.v, sf / RATE)


class C2M(C2):
    """C2 with every voice's cutoff driven by its own LFO (an emitter on LowPass.cutoff, fx.py:124-129): the filter is
    designed on the device once per request (k_design) and must stay on the same time-parallel kernel -- A/B against C2."""
    name = 'C2 with a modulated cutoff per voice (Mix / Sine LFO emitters on LowPass.cutoff), %d voices x %g s @ 48 kHz per GPU'
    kernel = 'k_param_eval + k_design (scan tables and decay horizon of the request) + k_chain_scan3'
    traffic_profile = None

    def build(self, ns):
        from signals_b200 import workloads as cases
        hertz, phase, cutoff, g = self.params
        rng = np.random.default_rng(1000 + self.rank)
        lfo_hz, lfo_ph = rng.uniform(0.1, 4.0, self.v), rng.uniform(0.0, 1.0, self.v)
        f = cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff])
        f.cutoff = cases.sweep(ns, [cutoff / 1.5], [cutoff * 1.5], [lfo_hz], [lfo_ph])
        return cases.gain(ns, f, [g])


class C3(Workload):
    name = 'C3: additive bank, %d sine partials -> %d channels (fused oscillator + mix reduction), %g s @ 48 kHz per GPU'
    kernel = 'k_bank (two transcendental points per eight samples + angle-addition rotations on the FMA pipe)'
    bound = 'sfu'
    bytes_per_unit = 4.0 / 1024

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        from signals_b200 import workloads as cases
        self.p = self.voices
        self.groups = max(1, self.p // 1024)
        self.out_channels = self.groups
        self.params = cases.bank_params(3 + self.rank, self.p, self.p // self.groups)

    def build(self, ns):
        from signals_b200 import workloads as cases
        from signals_b200.chain import ext
        return cases.build_bank(ns, ext, *self.params, self.groups)

    def describe(self):
        return {'workload': self.name % (self.p, self.groups, self.seconds), 'partials_per_gpu': self.p, 'frames': self.frames,
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
    kernel = 'k_cascade_delta (two channels per thread, all 8 sections in registers in delta form: 3 FFMA2 + 2 FADD2 per section, equal time pieces per warp slot)'
    traffic_profile = 'r02_k_cascade_delta_full.txt'
    bound = 'hbm'
    bytes_per_unit = 8.0

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.ch = self.voices
        self.out_channels = self.ch
        self.slab_frames = min(self.frames, int(self.slab_seconds * RATE))
        rng = np.random.default_rng(4 + self.rank)
        self.cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, self.ch)))
        self.seed = 4 + self.rank

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
        return {'workload': self.name % (self.ch, self.seconds, self.slab_frames / RATE), 'channels_per_gpu': self.ch, 'frames': self.frames, 'rate': RATE,
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
    kernel = 'k_voices (equal pieces of the (voice group, row block) space per CTA slot) + k_voices_finish'
    bound = 'sfu'
    bytes_per_unit = 0.0
    scaling = 'strong'
    out_channels = 2

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        from signals_b200 import workloads as cases
        self.n = self.voices
        self.prm = cases.instance_params(5, self.n, self.rank, self.world)

    def build(self, ns):
        from signals_b200 import workloads as cases
        from signals_b200.chain import ext
        return cases.build_instances(ns, ext, self.prm)

    def describe(self):
        return {'workload': self.name % (self.n, self.seconds), 'instances_total': self.n, 'frames': self.frames, 'rate': RATE,
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


WORKLOADS = {'c2': C2, 'c2m': C2M, 'c3': C3, 'c4': C4, 'c5': C5}


def make_workload(name, args, rank, world, headline):
    v, s, _, _ = DEFAULTS[name]
    if headline:
        v, s = args.voices, args.seconds
    return WORKLOADS[name](v, s, rank, world, args.slab_seconds)


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's numpy/scipy render (never on the product path)
# ------------------------------------------------------------------------------------------------
def _cpu_job(job):
    from oracle import np_oracle
    kind, a = job
    if kind == 'chain':
        hertz, phase, cutoff, g, frames = a
        out = np_oracle.render_voice_chain(0, frames, RATE, hertz, phase, cutoff, g)
    elif kind == 'blockwise':
        hertz, phase, cutoff, g, frames, block = a
        from signals_b200 import workloads as cases
        ns = cases.b200_namespace()      # node objects only: the oracle walks them with the reference's recursion, on the CPU
        graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])
        orc = np_oracle.GraphOracle(RATE)
        out = None
        for p in range(0, frames, block):
            out = orc.render(graph, p, min(block, frames - p), len(hertz))
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


def cpu_callback_latency(graphs):
    """C1 beside the GPU: the oracle port walking the same graphs with the reference's recursion, one request per
    callback block (per-block butter() per channel + 100-frame context, fx.py:93-105), 1 core."""
    from oracle import np_oracle
    rows = []
    for name, build in graphs:
        graph = build()
        orc = np_oracle.GraphOracle(RATE)
        for frames in (128, 512):
            lat = []
            pos = 0
            for _ in range(200):
                t0 = time.perf_counter_ns()
                orc.render(graph, pos, frames, 1)
                lat.append((time.perf_counter_ns() - t0) * 1e-3)
                pos += frames
            rows.append({'graph': name, 'frames': frames, 'p50_us': float(np.percentile(lat, 50)), 'p99_us': float(np.percentile(lat, 99))})
    return {'what': "oracle port walking the same graph with the reference's recursion, one request per block "
                    '(per-block butter() + 100-frame context, fx.py:93-105), 1 core', 'rows': rows}


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: the Python reference cannot travel
    to the GPU box), all host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    workers = os.cpu_count() or 1
    name = args.config if args.config in WORKLOADS else 'c2'
    wl = make_workload(name, args, 0, max(1, args.gpus), True)
    jobs, units, sample = wl.cpu_sample(workers)
    times = []
    _cpu_init()
    with mp.get_context('fork').Pool(workers, initializer=_cpu_init) as pool:     # pool start-up is not part of the steady state
        for step in range(args.warmup + args.steps):
            dt = cpu_time(jobs, workers, pool)
            if step >= args.warmup:
                times.append(dt)
        blockwise = None
        if hasattr(wl, 'cpu_sample_blockwise'):
            bjobs, bunits, bsample = wl.cpu_sample_blockwise(workers)
            bdt = cpu_time(bjobs, workers, pool)
            blockwise = {'value': bunits / bdt, 'unit': 'voice-samples/s', 'cores': workers, 'sample': bsample + ', %d processes' % workers,
                         'what': "the reference's product mode (SinkDevice callbacks of 512 frames, fx.py:93-105)"}
    total = sum(times)
    value = units * len(times) / total
    config = wl.describe()
    config['sampled'] = True
    config['sample'] = sample + '; throughput extrapolated linearly in voice-samples (voices are independent)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voice-samples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': wl.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config,
            'cpu_baseline': {'value': value, 'unit': 'voice-samples/s', 'cores': workers, 'kind': 'port',
                             'sample': sample + ', %d processes' % workers},
            'e2e': {'value': value, 'unit': 'voice-samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': 'numpy/scipy oracle port of the reference render (the Python reference cannot travel to the GPU box); '
                    'single-request mode, the favourable one for the CPU -- see cpu_baseline.blockwise_512 for its product mode'}
    if blockwise:
        line['cpu_baseline']['blockwise_512'] = blockwise
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


def gpu_locality(index):
    """NUMA node of the GPU and the CPUs next to it; the rank is pinned there BEFORE it allocates page-locked memory
    (first touch), so that the device->host copies of N ranks do not all cross one socket."""
    info = {'numa_node': None, 'cpus_bound': None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = '/sys/bus/pci/devices/%s/numa_node' % bus.lower()[-12:]
        if os.path.exists(path):
            info['numa_node'] = int(open(path).read().strip())
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info['cpus_bound'] = len(cpus)
    except Exception as e:   # noqa: BLE001
        info['error'] = type(e).__name__
    return info


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


class Ctx:
    """Per-process state shared by the measurements of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; the block render has no CPU fallback')
        self.locality = gpu_locality(self.local)
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        from signals_b200 import workloads as cases
        from signals_b200 import _lib, engine
        for kv in args.default_opt:
            k, val = kv.split('=')
            assert _lib.lib().sigb_set_default_option(k.encode(), int(val)) == 0, kv
        self.ns = cases.b200_namespace()
        self.eng = engine.Engine(device=torch.device('cuda', self.local))
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                self.peaks = json.load(f)
        except OSError:
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def configure(self, c):
        if self.args.scan_variant is not None:
            c.set_option('scan_variant', self.args.scan_variant)
        for kv in self.args.plan_opt:
            k, val = kv.split('=')
            c.set_option(k, int(val))

    def release(self):
        gc.collect()
        self.torch.cuda.empty_cache()


def measure(ctx, wl, steps, warmup, e2e_steps, cpu_baseline):
    """The bench contract for one workload: W warm-up steps, K steps timed with CUDA events on the launching stream
    between barriers, max over ranks; then the end-to-end leg through the public API with host buffers."""
    torch, dist = ctx.torch, ctx.dist
    from signals_b200 import shard
    world, rank = ctx.world, ctx.rank
    frames = wl.frames
    graph = wl.build(ctx.ns)
    t0 = time.perf_counter()
    compiled = ctx.eng.compile(graph, wl.out_channels, RATE, frames)
    compile_s = time.perf_counter() - t0
    ctx.configure(compiled)
    slab = wl.slab_frames or frames
    out = torch.empty((slab, wl.out_channels), dtype=torch.float32, device='cuda')
    reduce_mix = isinstance(wl, C5)

    def step(c):
        """One pass of the hot path over the whole workload, all on torch's current stream."""
        if reduce_mix:
            shard.render_reduced(c, 0, frames, out, dst=0)
            return
        for r in range(0, frames, slab):
            if wl.slab_frames:
                c.bind_window(wl.buffer, wl.noise, r)
            c.render_device(r, min(slab, frames - r), out)

    warm = max(warmup, 3)
    for _ in range(warm):
        step(compiled)
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    launches0 = compiled.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        step(compiled)
        ev[i + 1].record()
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[steps])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    last_ms = compiled.last_kernel_ms()          # device time of the last sigb_render call (library's own events)
    clocks = sampler.finish()
    ctx.barrier()
    launches = compiled.launch_count - launches0
    total_ms_max = ctx.max_over_ranks(total_ms)
    value = wl.units_per_step() * steps / (total_ms_max * 1e-3)
    res = {'value': value, 'ms_per_step': total_ms_max / steps, 'steps': steps, 'warmup': warm, 'launches': int(launches),
           'clocks': clocks, 'step_ms_min': float(np.min(step_ms)), 'step_ms_max': float(np.max(step_ms)),
           'compile_s': compile_s}
    res['checksum_last_frame'] = float(out[-1].double().sum())
    res['checksum_block'] = float(out.double().sum())

    if reduce_mix:
        # the collective alone (one (frames, 2) float32 reduce to rank 0), and the N-rank mix against ONE GPU's render of
        # the whole bank: rank 0 compiles all instances and renders them alone
        reps = 10
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            shard.reduce_mix(out, dst=0)
        e1.record()
        torch.cuda.synchronize()
        res['reduce_ms'] = ctx.max_over_ranks(e0.elapsed_time(e1) / reps)
        res['reduce_bytes'] = int(out.numel() * 4)
        step(compiled)                               # `out` on rank 0 = the N-rank mix again
        torch.cuda.synchronize()
        if rank == 0:
            res['checksum_last_frame'] = float(out[-1].double().sum())
            res['checksum_block'] = float(out.double().sum())
            if world > 1:
                whole = C5(wl.n, wl.seconds, 0, 1)
                c1 = ctx.eng.compile(whole.build(ctx.ns), 2, RATE, frames)
                ref = c1.render_device(0, frames)
                torch.cuda.synchronize()
                err = float((ref - out).abs().max())
                res['parity_n_vs_1'] = {'max_abs': err, 'tolerance': 1e-6, 'ok': bool(err <= 1e-6), 'mix_peak': float(ref.abs().max()),
                                        'checksum_last_frame_n1': float(ref[-1].double().sum()),
                                        'what': 'whole (frames, 2) mix of %d ranks vs rank 0 rendering all %d instances alone' % (world, wl.n)}
                c1.close()
                del ref, c1, whole
                assert err <= 1e-6, 'C5: %d-rank mix differs from the single-GPU mix by %.3e' % (world, err)
            else:
                res['parity_n_vs_1'] = {'max_abs': 0.0, 'tolerance': 1e-6, 'ok': True, 'what': 'N = 1: this IS the single-GPU mix'}

    # ---- end to end through the public API with HOST buffers: compile (host->device tables) +
    #      render_host (kernels + pipelined device->host copies), every step
    # streamed workloads: the host-buffer leg moves 2.5 s slabs (7.9 GB pinned each way for C4) -- it is bound by
    # the PCIe copies, and the pinned staging stays small next to the device-resident slabs of the timed leg
    e2e = None
    if e2e_steps > 0:
        eslab = min(slab, int(2.5 * RATE)) if wl.slab_frames else slab
        host_out = torch.empty((eslab, wl.out_channels), dtype=torch.float32, pin_memory=True)
        host_in = wl.noise[:eslab].cpu().pin_memory() if wl.slab_frames else None
        e2e_times, copy_wait = [], []
        h2d = int(compiled.describe()['param_bytes'])
        e2e_launches = 0
        E2E_WARMUP = 3                                  # untimed end-to-end steps first (pinned pages touched, PCIe link and copy engines warm)
        for i in range(e2e_steps + E2E_WARMUP):
            ctx.barrier()
            t0 = time.perf_counter()
            c2 = ctx.eng.compile(graph, wl.out_channels, RATE, frames)
            ctx.configure(c2)
            t1 = time.perf_counter()
            for r in range(0, frames, eslab):
                if wl.slab_frames:
                    dev_in = host_in.to('cuda', non_blocking=True)        # this slab's input: pinned host -> HBM, on torch's
                    c2.bind_window(wl.buffer, dev_in, r)                  # stream; render_host orders itself after it
                c2.render_host(r, min(eslab, frames - r), host_out)
            if reduce_mix and world > 1:
                mix = host_out.to('cuda', non_blocking=True)
                shard.reduce_mix(mix, dst=0)
                host_out.copy_(mix)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e_launches = c2.launch_count
            c2.close()
            if i >= E2E_WARMUP:
                e2e_times.append(dt)
                copy_wait.append(dt - (t1 - t0))
        if wl.slab_frames:
            h2d += int(host_in.numel() * 4 * ((frames + eslab - 1) // eslab))
        # what the platform gives a plain pinned device->host copy of one such block while all N ranks copy at once:
        # the ceiling of the host-buffer leg (at N = 8 the ranks share the host's PCIe uplinks)
        probe = []
        for _ in range(2):
            ctx.barrier()
            t0 = time.perf_counter()
            host_out.copy_(out[:eslab], non_blocking=True)
            torch.cuda.synchronize()
            probe.append(host_out.numel() * 4 / (time.perf_counter() - t0) / 1e9)
        print('bench: e2e step seconds %s (compile part %s)' % (['%.4f' % t for t in e2e_times], ['%.4f' % (t - w) for t, w in zip(e2e_times, copy_wait)]),
              file=sys.stderr, flush=True)
        te = ctx.max_over_ranks(sum(e2e_times))
        d2h = int(4 * wl.out_channels * frames)
        e2e = {'value': wl.units_per_step() * len(e2e_times) / te, 'unit': 'voice-samples/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'steps': len(e2e_times), 'warmup_steps': E2E_WARMUP,
               'what': 'Engine.compile(graph) + CompiledPlan.render_host(pinned fp32 block), every step',
               'step_s_min': float(np.min(e2e_times)), 'step_s_max': float(np.max(e2e_times)),
               'render_host_gbs_this_rank': (d2h + h2d) / float(np.mean(copy_wait)) / 1e9,
               'plain_d2h_memcpy_gbs_this_rank_all_ranks_copying': float(max(probe)),
               'compile_s_mean': float(np.mean(e2e_times) - np.mean(copy_wait)),
               'pinned_numa_node': ctx.locality.get('numa_node'), 'cpus_bound': ctx.locality.get('cpus_bound'),
               'gpu_launches_per_step': int(e2e_launches)}
        res['checksum_last_frame_e2e'] = float(host_out[-1].double().sum())
        del host_out, host_in

    # ---- roofline of the dominant kernel
    n_renders = (frames + slab - 1) // slab
    avg_step_ms = float(np.mean(step_ms))
    launch_ms = avg_step_ms / n_renders                       # one dominant-kernel launch per render call
    peaks = ctx.peaks
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
        if isinstance(wl, C4):
            # the co-limit: the FP32 pipe.  Delta-form sections (k_cascade_delta, default): 8 x (3 FFMA2 + 2 FADD2) per two
            # channel-samples = 40 lane-ops per channel-sample; state-variable sections (reg_variant 1 / 3 / 4): 8 x 6 = 48
            svf = any(o.startswith('reg_variant=') and o.split('=')[1] in ('1', '3', '4') for o in ctx.args.plan_opt)
            lane_ops = 48.0 if svf else 40.0
            sm_mhz = clocks.get('sm_mhz') or peaks.get('sm_max_mhz', 1965.0)
            fma = lane_ops * wl.launch_units() / (launch_ms * 1e-3) / (148 * sm_mhz * 1e6)
            roof['fma_lane_ops_per_clk_sm'] = fma
            roof['fma_lane_ops_per_channel_sample'] = lane_ops
            roof['fma_probe_ceiling'] = FMA_PROBE_CEILING
            roof['fma_frac_of_probe'] = fma / FMA_PROBE_CEILING
    else:
        # transcendental-bound kernels: one MUFU.SIN per unit on the 16-lane/clk/SM special-function pipe
        sm_mhz = clocks.get('sm_mhz') or peaks.get('sm_max_mhz', 1965.0)
        peak = 148 * 16 * sm_mhz * 1e6 / 1e9
        achieved = wl.launch_units() / (launch_ms * 1e-3) / 1e9
        roof = {'bound': 'sfu', 'achieved': achieved, 'peak': peak, 'unit': 'Gsample/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': '148 SMs x 16 MUFU lanes/clk x measured SM clock (derived; no MEASURED_PEAKS entry for the SFU pipe); '
                               'a kernel that also evaluates sines on the FMA pipe can exceed it',
                'algorithmic_bytes_per_voice_sample': wl.bytes_per_unit}
    roof['kernel'] = wl.kernel
    roof['launch_ms'] = launch_ms
    roof['last_render_ms_in_library'] = last_ms
    res['roofline'] = roof
    res['e2e'] = e2e
    res['config'] = wl.describe()
    res['scaling'] = wl.scaling
    if cpu_baseline and rank == 0:
        jobs, units, sample = wl.cpu_sample(1)
        _cpu_init()
        secs = cpu_time(jobs, 1)
        res['cpu_baseline'] = {'value': units / secs, 'unit': 'voice-samples/s', 'cores': 1, 'kind': 'port',
                               'sample': sample + ', 1 process (the reference is single-threaded); %.1f s of CPU' % secs,
                               'host_cores': os.cpu_count()}
        if hasattr(wl, 'cpu_sample_blockwise'):
            bjobs, bunits, bsample = wl.cpu_sample_blockwise(1)
            bsecs = cpu_time(bjobs, 1)
            res['cpu_baseline']['blockwise_512'] = {
                'value': bunits / bsecs, 'unit': 'voice-samples/s', 'cores': 1, 'sample': bsample + '; %.1f s of CPU' % bsecs,
                'what': "the reference's product mode (SinkDevice callbacks of 512 frames, fx.py:93-105)"}
    compiled.close()
    del out, compiled, graph
    if hasattr(wl, 'noise'):
        del wl.noise, wl.buffer
    ctx.release()
    return res


# ------------------------------------------------------------------------------------------------
# C1: the audio callback (latency)
# ------------------------------------------------------------------------------------------------
def c1_graphs(ns):
    """(name, build() -> emitter) of the two realtime graphs: scripts/example_sine.py as a graph, and the reference's
    lowpass_test.sigs fixture's audio path (Triangle 440 -> Gain 0.2 -> LowPass 600; the sink is attached to the LowPass,
    above the FileWriter / Wave taps, so that the number is the render, not the WAV file or the GUI queue)."""
    from signals_b200 import sigs
    from signals_b200 import workloads as cases

    def sine_gain():
        return cases.gain(ns, cases.osc(ns, 'Sine', [[500.0]]), [[0.2]])

    def lowpass_patch():
        patch = sigs.load(os.path.join(ROOT, 'tests', 'golden', 'lowpass_test.sigs'))
        (lp,) = [n for n in patch.nodes.values() if type(n).__name__ == 'LowPass']
        return lp

    return [('Sine<-Fixed -> Gain (scripts/example_sine.py as a graph)', sine_gain), ('lowpass_test.sigs (audio path)', lowpass_patch)]


def measure_c1(ctx, callbacks=1500, cpu=True):
    """p50 / p99 of one SinkDevice callback (dev.py:167-179) at 128 / 384 / 512 / 1024 frames: host wall clock around
    SinkDevice.render_block(outdata) -- plan look-up (graph epoch), ONE CUDA graph launch, stream sync, copy into the
    device's pageable float32 buffer."""
    from signals_b200.chain import dev
    info = dev.DeviceInfo(name='bench', index=0, hostapi=0, max_input_channels=0, max_output_channels=2,
                          default_low_input_latency=0.0, default_low_output_latency=0.0, default_high_input_latency=0.0,
                          default_high_output_latency=0.0, default_samplerate=float(RATE))
    rows = []
    for name, build in c1_graphs(ctx.ns):
        for frames in (128, 384, 512, 1024):
            for mode in ('graph', 'direct'):
                sink = dev.SinkDevice(info)
                sink.input = build()
                outdata = np.zeros((frames, 1), dtype=np.float32)
                ctx.eng.clear()
                from signals_b200 import engine as engine_mod
                engine_mod.default_engine().clear()
                sink.render_block(outdata, frames, RATE)             # compile + seek block
                compiled = engine_mod.default_engine().plan_for(sink._ports['input'].sig, 1, RATE, frames)
                compiled.set_option('rt_graph', 1 if mode == 'graph' else 0)
                for _ in range(50):
                    sink.render_block(outdata, frames, RATE)
                g0, l0 = compiled.graph_launches, compiled.launch_count
                lat = np.empty(callbacks)
                for i in range(callbacks):
                    t0 = time.perf_counter_ns()
                    sink.render_block(outdata, frames, RATE)
                    lat[i] = (time.perf_counter_ns() - t0) * 1e-3
                row = {'graph': name, 'frames': frames, 'mode': mode, 'callbacks': callbacks,
                       'p50_us': float(np.percentile(lat, 50)), 'p99_us': float(np.percentile(lat, 99)), 'max_us': float(lat.max()),
                       'block_period_us': frames / RATE * 1e6,
                       'cuda_graph_launches': compiled.graph_launches - g0, 'gpu_launches': compiled.launch_count - l0,
                       'checksum_last_block': float(np.abs(outdata).sum())}
                rows.append(row)
                engine_mod.default_engine().clear()
    res = {'what': 'SinkDevice.render_block(outdata): plan look-up + sigb_render_block (one captured CUDA graph launch per block, '
                   'position from a pinned block header, output in pinned staging) + copy into the pageable device buffer; '
                   'host wall clock per callback, blocks contiguous (carried filter state)',
           'rows': rows}
    if cpu:
        res['cpu_reference_blockwise'] = cpu_callback_latency(c1_graphs(ctx.ns))
    return res


def run_b200(args):
    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    if args.config == 'c1':
        res = measure_c1(ctx, cpu=not args.no_cpu_baseline) if rank == 0 else None
        if rank == 0:
            p = [r for r in res['rows'] if r['frames'] == 512 and r['mode'] == 'graph']
            line = {'metric': 'callback latency p99 @ 512 frames', 'value': max(r['p99_us'] for r in p), 'unit': 'us', 'n_gpus': world,
                    'steps': p[0]['callbacks'], 'warmup': 50, 'ms_per_step': float(np.mean([r['p50_us'] for r in p])) * 1e-3,
                    'higher_is_better': False, 'scaling': 'replicas only', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                    'config': {'workload': 'C1: audio callback blocks through SinkDevice.render_block'}, 'c1': res}
            print(json.dumps(line), flush=True)
        if world > 1:
            ctx.dist.destroy_process_group()
        return
    wl = make_workload(args.config, args, rank, world, True)
    head = measure(ctx, wl, args.steps, args.warmup, args.e2e_steps, cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    extra = {}
    if args.config == 'c2' and not args.no_extra and not args.voices_given:
        wanted = [x for x in args.extra.split(',') if x]
        for name in wanted:
            # a failure in an appended config (an out-of-memory slab, a parity assertion) is recorded under its key -- the
            # headline line is printed regardless
            try:
                if name == 'c5':                                   # every N: strong scaling with the NCCL reduce
                    r = measure(ctx, make_workload('c5', args, rank, world, False), EXTRA_STEPS['c5'], 3, 0, cpu_baseline=False)
                elif name in ('c3', 'c4', 'c2m') and world == 1:
                    r = measure(ctx, make_workload(name, args, rank, world, False), EXTRA_STEPS[name], EXTRA_WARMUP.get(name, 3), 0, cpu_baseline=False)
                elif name == 'c1' and world == 1:
                    r = measure_c1(ctx, cpu=not args.no_cpu_baseline)
                else:
                    continue
            except Exception as exc:      # noqa: BLE001
                extra[name] = {'error': '%s: %s' % (type(exc).__name__, str(exc)[:300])}
                print('bench: extra.%s failed: %r' % (name, exc), file=sys.stderr, flush=True)
                continue
            if name != 'c1':
                r = {k: r[k] for k in ('value', 'ms_per_step', 'steps', 'warmup', 'scaling', 'roofline', 'launches', 'clocks', 'config',
                                       'checksum_last_frame', 'checksum_block', 'compile_s', 'reduce_ms', 'reduce_bytes', 'parity_n_vs_1')
                     if k in r}
                r['unit'] = 'voice-samples/s'
                r['n_gpus'] = world
                if name == 'c2m':
                    r['vs_unmodulated_c2'] = r['value'] / head['value']
            extra[name] = r
    if rank == 0:
        line = {'metric': METRIC, 'value': head['value'], 'unit': 'voice-samples/s', 'n_gpus': world, 'steps': head['steps'],
                'warmup': head['warmup'], 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
                'scaling': head['scaling'], 'vs_baseline': None, 'dtype': 'f32 (fp64 phase / Q0.64 phase accumulator, fp64 scan carries)',
                'data': 'synthetic', 'config': head['config'], 'roofline': head['roofline'], 'e2e': head['e2e'],
                'gpu_launches': head['launches'], 'clocks': head['clocks'], 'step_ms_min': head['step_ms_min'],
                'step_ms_max': head['step_ms_max'], 'checksum_last_frame': head.get('checksum_last_frame_e2e', head['checksum_last_frame']),
                'checksum_block': head['checksum_block'], 'compile_s': head['compile_s']}
        for k in ('reduce_ms', 'reduce_bytes', 'parity_n_vs_1', 'cpu_baseline'):
            if k in head:
                line[k] = head[k]
        if extra:
            line['extra'] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
