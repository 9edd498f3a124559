"""Synthetic workloads of the BASELINE configs: parameter distributions and graph builders.

Used by ``bench.py`` to build what it measures and -- re-exported by ``oracle/cases.py`` -- by the parity tests, so
that the benchmark, the tests and the CPU baseline all draw the very same voices.  The builders take a *namespace* of
node classes (``b200_namespace()`` = the product's mirror of the reference API), so the same graph can also be built
from the unmodified reference in the build container.
"""
from __future__ import annotations

import types

import numpy as np

RATE = 48000


def b200_namespace() -> types.SimpleNamespace:
    import signals_b200.chain.fixed as fixed
    import signals_b200.chain.fx as fx
    import signals_b200.chain.osc as osc
    import signals_b200.chain.shape as shape
    return types.SimpleNamespace(
        Fixed=fixed.Fixed,
        Sine=osc.Sine, Square=osc.Square, Sawtooth=osc.Sawtooth, Triangle=osc.Triangle,
        Mix=fx.Mix, RingMod=fx.RingMod, Gain=fx.Gain, Amp=fx.Amp,
        LowPass=fx.LowPass, HighPass=fx.HighPass,
        Merge=shape.Merge,
    )


def fixed(ns, value, enabled: bool = True):
    f = ns.Fixed()
    f.get_state().value = np.array(value, ndmin=2, dtype=float)
    if not enabled:
        f.get_state().enabled = False
    return f


def osc(ns, wave: str, hertz, phase=None):
    o = getattr(ns, wave)()
    o.hertz = fixed(ns, hertz)
    if phase is not None:
        o.phase = fixed(ns, phase)
    return o


def gain(ns, left, right):
    g = ns.Gain()
    g.left = left
    g.right = fixed(ns, right)
    return g


def lowpass(ns, input_, cutoff, cls: str = 'LowPass', order: int | None = None):
    base = getattr(ns, cls)
    if order is not None:
        # the reference's class-level knob, chain/fx.py:66
        base = type(f'{cls}{order}', (base,), {'order': order})
    f = base()
    f.input = input_
    f.cutoff = fixed(ns, cutoff)
    return f


def sweep(ns, lo, hi, lfo_hertz, lfo_phase):
    """A block-rate parameter sweeping [lo, hi] (arrays broadcast per channel): Mix(hi, lo, mix = 0.5 + 0.25 * sine LFO),
    an emitter for a filter's cutoff port, sampled once per request (SingleCritFilter._eval, chain/fx.py:124-129)."""
    lfo = osc(ns, 'Sine', lfo_hertz, lfo_phase)
    half = ns.Mix()
    half.left = fixed(ns, [[1.0]])
    half.right = fixed(ns, [[0.0]])
    half.mix = gain(ns, lfo, [[0.5]])      # in [-0.5, 0.5]
    w = ns.Mix()                           # 0.5 * 1 + 0.5 * half, i.e. [0.25, 0.75]
    w.left = fixed(ns, [[1.0]])
    w.right = half
    w.mix = fixed(ns, [[0.5]])
    m = ns.Mix()
    m.left = fixed(ns, hi)
    m.right = fixed(ns, lo)
    m.mix = w
    return m


def voice_params(seed: int, v: int):
    """SURVEY 8d / BASELINE config C2 parameter distributions."""
    rng = np.random.default_rng(seed)
    hertz = rng.uniform(27.5, 4186.0, v)
    phase = rng.uniform(0.0, 1.0, v)
    cutoff = np.exp(rng.uniform(np.log(100.0), np.log(8000.0), v))
    g = rng.uniform(0.05, 1.0, v)
    return hertz, phase, cutoff, g


def bank_params(seed: int, partials: int, per_group: int = 1024):
    """C3: hertz~U(27.5, 12000), phase~U(0,1), amp~U(0,1)/per_group; group g = p // per_group."""
    rng = np.random.default_rng(seed)
    hertz = rng.uniform(27.5, 12000.0, partials)
    phase = rng.uniform(0.0, 1.0, partials)
    amp = rng.uniform(0.0, 1.0, partials) / per_group
    return hertz, phase, amp


def build_bank(ns, ext, hertz, phase, amp, groups: int):
    """GroupSum(Gain(Sine(hertz, phase), amp)) -> (frames, groups)."""
    gs = ext.GroupSum()
    gs.get_state().groups = groups
    gs.input = gain(ns, osc(ns, 'Sine', [hertz], [phase]), [amp])
    return gs


WAVES = ('Sine', 'Square', 'Sawtooth', 'Triangle')
FILTERS = (None, 'LowPass', 'HighPass')


def instance_params(seed: int, n_total: int, rank: int = 0, world: int = 1) -> dict:
    """C5: n_total randomised osc -> (filter) -> gain -> pan instances; instance i lives on rank
    i % world.  Returns this rank's shard as arrays (global draw, so every world size sees the same bank)."""
    rng = np.random.default_rng(seed)
    wave = rng.integers(0, 4, n_total)
    filt = rng.integers(0, 3, n_total)
    hertz = rng.uniform(27.5, 4186.0, n_total)
    phase = rng.uniform(0.0, 1.0, n_total)
    cutoff = np.exp(rng.uniform(np.log(100.0), np.log(8000.0), n_total))
    g = rng.uniform(0.05, 1.0, n_total) / np.sqrt(n_total)
    pan = rng.uniform(0.0, 1.0, n_total)
    mine = slice(rank, n_total, world)
    return dict(wave=wave[mine], filt=filt[mine], hertz=hertz[mine], phase=phase[mine], cutoff=cutoff[mine],
                gain=g[mine], pan=pan[mine], n_total=n_total)


def build_instances(ns, ext, prm: dict):
    """One homogeneous chain per (wave, filter) kind, merged channel-wise, under one PanSum."""
    chains, pans = [], []
    for w, wname in enumerate(WAVES):
        for f, fname in enumerate(FILTERS):
            sel = np.flatnonzero((prm['wave'] == w) & (prm['filt'] == f))
            if sel.size == 0:
                continue
            x = osc(ns, wname, [prm['hertz'][sel]], [prm['phase'][sel]])
            if fname is not None:
                x = lowpass(ns, x, [prm['cutoff'][sel]], fname)
            chains.append(gain(ns, x, [prm['gain'][sel]]))
            pans.append(prm['pan'][sel])
    node = chains[0]
    for nxt in chains[1:]:
        m = ns.Merge()
        m.left = node
        m.right = nxt
        node = m
    ps = ext.PanSum()
    ps.input = node
    ps.pan = fixed(ns, [np.concatenate(pans)])
    return ps
