#!/usr/bin/env python
"""Ad-hoc timing of C2's chain (sine -> low-pass -> gain, 4,096 voices x 10 s) with the oscillator's hertz driven by an LFO
(vibrato: a block-rate parameter, sampled once per request) against the same chain with constant hertz."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from signals_b200 import workloads as cases   # noqa: E402
from signals_b200 import engine   # noqa: E402

RATE, CH, FRAMES = 48000, 4096, 480000
ns = cases.b200_namespace()
hz, ph, cut, g = cases.voice_params(2, CH)
rng = np.random.default_rng(3)
out = torch.empty((FRAMES, CH), dtype=torch.float32, device='cuda')
WAVE = os.environ.get('WAVE', 'Sine')
for nsec in (0, 1, 2):
    for mod in (False, True) if nsec else (False, 'seq', True):
        o = getattr(ns, WAVE)()
        o.hertz = cases.sweep(ns, [hz * 0.97], [hz * 1.03], [rng.uniform(3.0, 7.0, CH)], [rng.uniform(0, 1, CH)]) if mod is True else cases.fixed(ns, [hz])
        o.phase = cases.fixed(ns, [ph])
        node = o
        for s in range(nsec):
            node = cases.lowpass(ns, node, [cut * (1.0 + 0.1 * s)])
        node = cases.gain(ns, node, [g])
        c = engine.Engine().compile(node, CH, RATE)
        if mod == 'seq':
            c.set_option('osc_fill', 0)
        for _ in range(3):
            c.render_device(0, FRAMES, out)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            c.render_device(0, FRAMES, out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print(f'{WAVE}: {nsec} section(s), hertz {"LFO-driven" if mod is True else "constant, k_chain_seq" if mod == "seq" else "constant"}: {ms:.3f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gvoice-samples/s, launches {[l["kind"] for l in c.describe()["launches"]]}')
        c.close()
