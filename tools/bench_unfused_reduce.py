#!/usr/bin/env python
"""Ad-hoc timing of reductions that do NOT fuse with their producers: GroupSum over filtered sine partials (C3 with a low-pass per
partial) and PanSum with the fusion switched off -- chain launches into a materialised block + k_reduce."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from signals_b200 import workloads as cases   # noqa: E402
from signals_b200 import engine   # noqa: E402
from signals_b200.chain import ext   # noqa: E402

RATE = 48000
ns = cases.b200_namespace()
P, G, FRAMES = 65536, 64, 96000
hertz, phase, amp = cases.bank_params(3, P, P // G)
cut = np.exp(np.random.default_rng(4).uniform(np.log(300.0), np.log(8000.0), P))
gs = ext.GroupSum()
gs.get_state().groups = G
gs.input = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cut]), [amp])
out = torch.empty((FRAMES, G), dtype=torch.float32, device='cuda')
c = engine.Engine().compile(gs, G, RATE, FRAMES)
for _ in range(2):
    c.render_device(0, FRAMES, out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(3):
    c.render_device(0, FRAMES, out)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 3
print(f'GroupSum(Gain(LowPass(Sine))) {P} partials -> {G} channels x {FRAMES / RATE:g} s: {ms:.2f} ms per render, {P * FRAMES / ms / 1e9 * 1e3:.4g} Gpartial-samples/s, '
      f'launches {[l["kind"] for l in c.describe()["launches"]]}, CUDA launches per render {c.launch_count // 5}')
c.close()
