#!/usr/bin/env python
"""Ad-hoc timing of the pointwise nodes on C2's shape (4,096 channels x 10 s): Mix / RingMod of two oscillators (fused as the epilogue of
a stateless chain), Amp, Mix of two filtered chains (materialised blocks)."""
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
hz2 = hz * 1.01
out = torch.empty((FRAMES, CH), dtype=torch.float32, device='cuda')


def mix(a, b, m=0.3):
    n = ns.Mix(); n.left = a; n.right = b; n.mix = cases.fixed(ns, [[m]]); return n


def ring(a, b):
    n = ns.RingMod(); n.left = a; n.right = b; return n


def amp(a, e=2.0):
    n = ns.Amp(); n.left = a; n.right = cases.fixed(ns, [[e]]); return n


graphs = {
    'Mix(Sine, Sine)': lambda: mix(cases.osc(ns, 'Sine', [hz], [ph]), cases.osc(ns, 'Sine', [hz2], [ph])),
    'RingMod(Sine, Square)': lambda: ring(cases.osc(ns, 'Sine', [hz], [ph]), cases.osc(ns, 'Square', [hz2], [ph])),
    'Mix(Gain(Sine), Gain(Sawtooth))': lambda: mix(cases.gain(ns, cases.osc(ns, 'Sine', [hz], [ph]), [g]), cases.gain(ns, cases.osc(ns, 'Sawtooth', [hz2], [ph]), [g])),
    'Amp(Sine, 2)': lambda: amp(cases.osc(ns, 'Sine', [hz], [ph])),
    'Mix(LP(Sine), LP(Sine))': lambda: mix(cases.lowpass(ns, cases.osc(ns, 'Sine', [hz], [ph]), [cut]), cases.lowpass(ns, cases.osc(ns, 'Sine', [hz2], [ph]), [cut])),
    'LP(Mix(Sine, Sine))': lambda: cases.lowpass(ns, mix(cases.osc(ns, 'Sine', [hz], [ph]), cases.osc(ns, 'Sine', [hz2], [ph])), [cut]),
}
for name, build in graphs.items():
    c = engine.Engine().compile(build(), CH, RATE)
    for _ in range(2):
        c.render_device(0, FRAMES, out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(4):
        c.render_device(0, FRAMES, out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 4
    print(f'{name}: {ms:.3f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gchannel-samples/s, launches {[l["kind"] + ":" + str(l.get("op", l.get("sections", ""))) for l in c.describe()["launches"]]}')
    c.close()

# Merge trees (np.hstack, shape.py:60-74): four 1,024-channel chains merged pairwise into the 4,096-channel block
def merge(a, b):
    m = ns.Merge(); m.left = a; m.right = b; return m


q = CH // 4
parts = [cases.gain(ns, cases.lowpass(ns, cases.osc(ns, w, [hz[k * q:(k + 1) * q]], [ph[k * q:(k + 1) * q]]), [cut[k * q:(k + 1) * q]]), [g[k * q:(k + 1) * q]])
         for k, w in enumerate(('Sine', 'Square', 'Sawtooth', 'Triangle'))]
node = merge(merge(merge(parts[0], parts[1]), parts[2]), parts[3])
c = engine.Engine().compile(node, CH, RATE)
for _ in range(2):
    c.render_device(0, FRAMES, out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(4):
    c.render_device(0, FRAMES, out)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 4
print(f'Merge tree of four osc -> LP -> gain chains: {ms:.3f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gchannel-samples/s, launches {[l["kind"] for l in c.describe()["launches"]]}')
c.close()

# tremolo: a Gain driven by an LFO at the end of C2's chain
trem = ns.Gain()
trem.left = cases.lowpass(ns, cases.osc(ns, 'Sine', [hz], [ph]), [cut])
trem.right = cases.sweep(ns, [np.full(CH, 0.2)], [np.full(CH, 1.0)], [np.full(CH, 5.0)], [ph])
c = engine.Engine().compile(trem, CH, RATE)
for _ in range(2):
    c.render_device(0, FRAMES, out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(4):
    c.render_device(0, FRAMES, out)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 4
print(f'Gain(LP(Sine), LFO) [tremolo]: {ms:.3f} ms per render, {CH * FRAMES / ms / 1e9 * 1e3:.4g} Gchannel-samples/s, launches {[l["kind"] for l in c.describe()["launches"]]}')
c.close()
