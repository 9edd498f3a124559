#!/usr/bin/env python
"""Ad-hoc timing of sine -> N low-pass sections -> gain on C2's shape (4,096 voices x 10 s, write-only 4 B per voice-sample) for
N = 1, 2, 3, 4: the plan's default kernel against k_osc_reg forced from N sections (osc_reg = N) and never (osc_reg = 0)."""
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
for nsec in (1, 2, 3, 4):
    node = cases.osc(ns, 'Sine', [hz], [ph])
    for s in range(nsec):
        node = cases.lowpass(ns, node, [cut * (1.0 + 0.1 * s)])
    node = cases.gain(ns, node, [g])
    variants = [('default', {}), ('osc_reg forced', {'osc_reg': nsec}), ('osc_reg forced, state-variable sections', {'osc_reg': nsec, 'osc_delta': 0}),
                ('osc_reg never', {'osc_reg': 0})]
    if len(sys.argv) > 1:        # sweep the time pieces (percent of the resident warp slots) of the register kernels
        variants = [('delta, pieces %s%%' % pct, {'osc_reg': nsec, 'osc_pieces_pct': int(pct)}) for pct in sys.argv[1:]]
        variants += [('state-variable, pieces %s%%' % pct, {'osc_reg': nsec, 'osc_delta': 0, 'osc_pieces_pct': int(pct)}) for pct in sys.argv[1:]]
    for name, opts in variants:
        c = engine.Engine().compile(node, CH, RATE)
        for k, v in opts.items():
            c.set_option(k, v)
        for _ in range(3):
            c.render_device(0, FRAMES, out)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            c.render_device(0, FRAMES, out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        print(f'{nsec} section(s), {name}: {ms:.3f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gvoice-samples/s, {4 * CH * FRAMES / ms / 1e6:.0f} GB/s written')
        c.close()
