#!/usr/bin/env python
"""Ad-hoc timing of 8-section cascades on a materialised block (16,384 channels x 10 s, 4 B read + 4 B written per
channel-sample): all-low-pass and all-high-pass, k_cascade_delta (reg_variant 0) against k_cascade_reg's state-variable
sections (reg_variant 4), and a mixed cascade: k_cascade_delta's select-per-section form against k_cascade_pipe (cascade_reg 0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from signals_b200 import workloads as cases   # noqa: E402
from signals_b200 import engine   # noqa: E402
from signals_b200.chain import ext   # noqa: E402

RATE, CH, FRAMES, NSEC = 48000, 16384, 480000, int(os.environ.get('NSEC', 8))
PCTS = [int(v) for v in os.environ.get('PCTS', '100').split(',')]
rng = np.random.default_rng(7)
ns = cases.b200_namespace()
g = torch.Generator(device='cuda')
g.manual_seed(7)
noise = torch.rand((FRAMES, CH), generator=g, device='cuda', dtype=torch.float32) * 2 - 1
cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (NSEC, CH)))
out = torch.empty((FRAMES, CH), dtype=torch.float32, device='cuda')
for kinds in ('L' * NSEC, 'H' * NSEC, 'H' * (NSEC // 2) + 'L' * (NSEC - NSEC // 2)):
    node = ext.Buffer(noise)
    for s, k in enumerate(kinds):
        node = cases.lowpass(ns, node, [cut[s]], 'HighPass' if k == 'H' else 'LowPass')
    mixed = len(set(kinds)) > 1
    for variant, pct in [(0, p) for p in PCTS] + [(4, 100)]:
        c = engine.Engine().compile(node, CH, RATE, FRAMES)
        c.set_option('reg_variant', variant)
        c.set_option('osc_pieces_pct', pct)
        if (mixed or NSEC == 2) and variant == 4:
            c.set_option('cascade_reg', 0)
        for _ in range(2):
            c.render_device(0, FRAMES, out)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            c.render_device(0, FRAMES, out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        other = 'k_chain_scan2' if NSEC == 2 else 'k_cascade_pipe' if mixed else 'k_cascade_reg (state-variable)'
        name = ('k_cascade_delta (mixed)' if mixed else 'k_cascade_delta') if variant == 0 else other
        print(f'{kinds} {name} pieces {pct}%: {ms:.2f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gchannel-samples/s, {8 * CH * FRAMES / ms / 1e6:.0f} GB/s read + written')
        c.close()
