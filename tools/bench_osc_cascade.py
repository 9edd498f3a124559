#!/usr/bin/env python
"""Ad-hoc timing of an oscillator -> 8 low-pass sections chain (16,384 channels x 10 s, write-only 4 B per
channel-sample): k_osc_reg (register-resident, default) against k_cascade_pipe (osc_reg=0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from signals_b200 import workloads as cases   # noqa: E402
from signals_b200 import engine   # noqa: E402

RATE, CH, FRAMES, NSEC = 48000, 16384, 480000, 8
rng = np.random.default_rng(7)
ns = cases.b200_namespace()
node = cases.osc(ns, 'Sine', [rng.uniform(27.5, 4186.0, CH)], [rng.uniform(0, 1, CH)])
cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (NSEC, CH)))
for s in range(NSEC):
    node = cases.lowpass(ns, node, [cut[s]])
out = torch.empty((FRAMES, CH), dtype=torch.float32, device='cuda')
for name, opts in (('k_osc_reg', {}), ('k_cascade_pipe', {'osc_reg': 0})):
    c = engine.Engine().compile(node, CH, RATE)
    for k, v in opts.items():
        c.set_option(k, v)
    for _ in range(2):
        c.render_device(0, FRAMES, out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        c.render_device(0, FRAMES, out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    print(f'{name}: {ms:.2f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gchannel-samples/s, {4 * CH * FRAMES / ms / 1e6:.0f} GB/s written')
    c.close()
