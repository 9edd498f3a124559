"""TEST INFRASTRUCTURE ONLY -- shared parity cases.

Each case builds a node graph from a *namespace* of node classes, so that the very same
graph can be built from the unmodified reference (``ref_namespace``; build container
only), and from ``signals_b200.chain`` (the product's mirror of the reference API).
``oracle/make_golden.py`` renders every case with the reference and commits the result
under ``tests/golden/``; ``tests/`` then checks the numpy oracle (CPU) and the CUDA path
(GPU) against those files.
"""
from __future__ import annotations

import dataclasses
import types
import typing

import numpy as np

from signals_b200.workloads import (RATE, FILTERS, WAVES, b200_namespace, bank_params, build_bank, build_instances, fixed, gain,   # noqa: F401
                                    instance_params, lowpass, osc, sweep, voice_params)


def ref_namespace(ref) -> types.SimpleNamespace:
    return types.SimpleNamespace(
        Fixed=ref.fixed.Fixed,
        Sine=ref.osc.Sine, Square=ref.osc.Square, Sawtooth=ref.osc.Sawtooth, Triangle=ref.osc.Triangle,
        Mix=ref.fx.Mix, RingMod=ref.fx.RingMod, Gain=ref.fx.Gain, Amp=ref.fx.Amp,
        LowPass=ref.fx.LowPass, HighPass=ref.fx.HighPass,
        Merge=ref.shape.Merge,
    )


@dataclasses.dataclass(frozen=True)
class Case:
    name: str
    build: typing.Callable[[typing.Any], typing.Any]
    frames: int
    channels: int
    position: int = 0
    rate: int = RATE
    tol: float = 1e-6
    # 'exact-edges': discontinuous waveform; samples must agree except none (bit-faithful fp64 phase)
    stride: int = 1          # golden stores out[::stride] (long renders)
    note: str = ''
    block: int = 0           # > 0: rendered as consecutive requests of `block` frames (block-rate parameters are
                             # re-sampled at the first frame of every request, chain/__init__.py:305-306)
    options: dict = dataclasses.field(default_factory=dict)     # sigb_plan_set_option settings the GPU render needs


def _c2(ns, v=8, seed=2, wave='Sine', cls='LowPass'):
    hertz, phase, cutoff, g = voice_params(seed, v)
    return gain(ns, lowpass(ns, osc(ns, wave, [hertz], [phase]), [cutoff], cls), [g])


def _lowpass_test_sigs(ns):
    """src/signals/lowpass_test.sigs minus the FileWriter/Wave/sink taps: Merge(left=LowPass, right=Gain)."""
    tri = osc(ns, 'Triangle', [[440]])
    g = gain(ns, tri, [[0.2]])
    lp = lowpass(ns, g, [[600]])
    m = ns.Merge()
    m.left = lp
    m.right = g
    return m


def _mix(ns):
    m = ns.Mix()
    m.left = osc(ns, 'Sine', [[220.0, 330.0, 440.0]], [[0.0, 0.25, 0.5]])
    m.right = osc(ns, 'Sawtooth', [[110.5, 221.25, 331.125]])
    m.mix = fixed(ns, [[0.25, 0.5, 0.9]])
    return m


def _ringmod(ns):
    r = ns.RingMod()
    r.left = osc(ns, 'Sine', [[441.0, 882.5]])
    r.right = osc(ns, 'Triangle', [[3.3, 7.7]], [[0.1, 0.2]])
    return r


def _amp(ns, exp):
    a = ns.Amp()
    a.left = osc(ns, 'Sine', [[100.0, 250.0]], [[0.05, 0.3]])
    a.right = fixed(ns, exp)
    return a


def _cascade8(ns, v=4, seed=4):
    rng = np.random.default_rng(seed)
    hertz = rng.uniform(55.0, 880.0, v)
    x = osc(ns, 'Sawtooth', [hertz], [rng.uniform(0, 1, v)])
    for _ in range(8):
        x = lowpass(ns, x, [np.exp(rng.uniform(np.log(200.0), np.log(8000.0), v))])
    return x


def _cascade2(ns):
    node = osc(ns, 'Sawtooth', [[220.0, 331.0]])
    for cut in ([[900.0, 2500.0]], [[1200.0, 3000.0]]):
        node = lowpass(ns, node, cut)
    return node


def _broadcast(ns):
    # 1-channel oscillator through a 4-channel Gain: (F,1)*(1,4) broadcast, chain/fx.py:52
    return gain(ns, osc(ns, 'Sine', [[330.0]]), [[0.1, 0.2, 0.3, 0.4]])


def _disabled_osc(ns):
    o = osc(ns, 'Sine', [[330.0, 331.0]])
    o.get_state().enabled = False
    return gain(ns, o, [[0.5, 0.5]])


def _unconnected(ns):
    # Mix with nothing on `right` and `mix`: zeros((1,1)) for both (chain/__init__.py:296-298)
    m = ns.Mix()
    m.left = osc(ns, 'Sine', [[330.0, 660.0]])
    return m


def _disabled_param(ns):
    o = ns.Sine()
    o.hertz = fixed(ns, [[500.0, 600.0]])
    o.phase = fixed(ns, [[0.25, 0.5]], enabled=False)   # disabled Fixed -> zeros(1,1)
    return o


def _fanout(ns):
    # one oscillator feeding both sides of a RingMod and a Merge (fan-out; the cache's job, chain/__init__.py:424)
    o = osc(ns, 'Sine', [[200.0, 300.0]])
    r = ns.RingMod()
    r.left = o
    r.right = gain(ns, o, [[0.5, 0.25]])
    m = ns.Merge()
    m.left = r
    m.right = lowpass(ns, o, [[1000.0, 2000.0]], 'HighPass')
    return m


def _lfo(ns, wave, hertz, phase=None):
    return osc(ns, wave, hertz, phase)


def _lfo_gain(ns):
    # tremolo: Gain.right driven by a 2.5 Hz sine (two channels, different LFO phases)
    g = ns.Gain()
    g.left = osc(ns, 'Sine', [[440.0, 661.5]])
    g.right = _lfo(ns, 'Sine', [[2.5, 3.25]], [[0.1, 0.6]])
    return g


def _lfo_hertz(ns):
    # vibrato: hertz = Mix(880, 220, mix = sawtooth LFO) -- the oscillator's frequency is sampled per request
    m = ns.Mix()
    m.left = fixed(ns, [[880.0, 660.0]])
    m.right = fixed(ns, [[220.0, 330.0]])
    m.mix = _lfo(ns, 'Sawtooth', [[0.7]], [[0.2]])
    o = ns.Sine()
    o.hertz = m
    o.phase = gain(ns, _lfo(ns, 'Triangle', [[1.3, 0.9]]), [[0.25, 0.5]])
    return o


def _lfo_mix_amp(ns):
    # crossfade driven by a square LFO scaled into [0.25, 0.75], then Amp with a modulated exponent
    m = ns.Mix()
    m.left = osc(ns, 'Sine', [[300.0, 450.0]])
    m.right = osc(ns, 'Triangle', [[200.0, 150.0]])
    mx = ns.Mix()
    mx.left = fixed(ns, [[0.75]])
    mx.right = fixed(ns, [[0.25]])
    mx.mix = gain(ns, _lfo(ns, 'Square', [[1.0]], [[0.1]]), [[0.5]])      # -> 0.5 +- ... stays a valid crossfade weight
    m.mix = mx
    a = ns.Amp()
    a.left = m
    ex = ns.Mix()                       # exponent = Mix(3, 2, mix = square LFO in {+1, -1}) in {3, 1}: integer-valued, so
    ex.left = fixed(ns, [[3.0]])        # negative inputs stay finite (a fractional exponent is NaN there, fx.py:60)
    ex.right = fixed(ns, [[2.0]])
    ex.mix = _lfo(ns, 'Square', [[1.0]], [[0.1]])
    a.right = ex
    return a


def _lfo_chain(ns):
    # modulated gain in front of a filter: the Gain cannot fold into the chain, the filter still fuses with its tail
    g = ns.Gain()
    g.left = osc(ns, 'Sawtooth', [[220.0, 331.0]])
    g.right = _lfo(ns, 'Sine', [[4.0]], [[0.3]])
    return gain(ns, lowpass(ns, g, [[900.0, 2500.0]]), [[0.5, 0.25]])


def _wah(ns, lo, hi, lfo_hertz, lfo_phase):
    # cutoff = Mix(hi, lo, mix = 0.5 + 0.25 * sine LFO) sweeps inside [lo, hi]: an emitter on the filter's cutoff port,
    # sampled once per request (SingleCritFilter._eval, fx.py:124-129)
    return sweep(ns, lo, hi, lfo_hertz, lfo_phase)


def _with_cutoff(ns, input_, cutoff_emitter, cls='LowPass', order=None):
    f = lowpass(ns, input_, [[1000.0]], cls, order)
    f.cutoff = cutoff_emitter
    return f


def _lfo_cutoff(ns):
    src = osc(ns, 'Sawtooth', [[220.0, 331.0]])
    return gain(ns, _with_cutoff(ns, src, _wah(ns, [[400.0, 900.0]], [[3000.0, 5200.0]], [[0.7, 1.1]], [[0.1, 0.4]])), [[0.5, 0.25]])


def _lfo_cutoff_hp3(ns):
    src = osc(ns, 'Square', [[220.5, 331.0]])
    return _with_cutoff(ns, src, _wah(ns, [[300.0, 700.0]], [[2500.0, 4000.0]], [[1.3, 0.9]], [[0.2, 0.7]]), 'HighPass', 3)


def _lfo_cutoff_cascade(ns):
    node = osc(ns, 'Sawtooth', [[110.0, 196.0]])
    for k in range(4):      # four chained LowPass nodes, each with its own sweeping cutoff: one 4-section launch
        node = _with_cutoff(ns, node, _wah(ns, [[600.0 + 150 * k, 900.0 + 100 * k]], [[4000.0 + 500 * k, 6000.0 - 300 * k]],
                                          [[0.5 + 0.2 * k, 0.8 + 0.1 * k]], [[0.1 * k, 0.3 + 0.1 * k]]))
    return node


CASES: list[Case] = [
    Case('sine_basic', lambda ns: osc(ns, 'Sine', [[440.0, 1000.0, 27.5, 4186.0]], [[0.0, 0.1, 0.5, 0.9]]), 4800, 4),
    Case('sine_pos1', lambda ns: osc(ns, 'Sine', [[440.0, 12000.0]], [[0.0, 0.37]]), 1000, 2, position=1),
    Case('sine_pos47999', lambda ns: osc(ns, 'Sine', [[440.0, 12000.0]], [[0.0, 0.37]]), 1000, 2, position=47999),
    Case('sine_pos2e31', lambda ns: osc(ns, 'Sine', [[440.0, 439.99]], [[0.0, 0.37]]), 1000, 2, position=2 ** 31 + 5,
         note='reference fp64 phase itself carries ~3e-8 cycles of rounding at n=2^31 (achieved 4.0e-7)'),
    Case('sine_60s_tail', lambda ns: osc(ns, 'Sine', [[4186.0, 27.5, 999.999]], [[0.0, 0.5, 0.123]]), 4800, 3,
         position=60 * RATE - 4800),
    Case('vis_test_sigs', lambda ns: osc(ns, 'Sine', [[220]]), 4800, 1,
         note='src/signals/vis_test.sigs: Sine 220 -> (Wave tap) -> sink'),
    Case('example_sine', lambda ns: gain(ns, osc(ns, 'Sine', [[500.0]]), [[0.2]]), 4800, 1,
         note='graph form of scripts/example_sine.py:50-53 (f=500, a=0.2)'),
    Case('edited_plot_sine330', lambda ns: osc(ns, 'Sine', [[330.0]]), 4608, 1,
         note='scripts/edited_plot.py:23-26,39-40'),
    Case('square_irrational', lambda ns: osc(ns, 'Square', [[439.99, 1234.567, 27.5001]], [[0.0, 0.3, 0.7]]), 4800, 3),
    Case('square_edges', lambda ns: osc(ns, 'Square', [[440.0, 1000.0, 12000.0, 6000.0]]), 48000, 4,
         note='rational hertz/rate: samples land exactly on the edge (sign(0)=0), SURVEY H1'),
    Case('sawtooth_irrational', lambda ns: osc(ns, 'Sawtooth', [[439.99, 1234.567]], [[0.0, 0.3]]), 4800, 2),
    Case('sawtooth_edges', lambda ns: osc(ns, 'Sawtooth', [[440.0, 1000.0, 12000.0]]), 48000, 3),
    Case('triangle_irrational', lambda ns: osc(ns, 'Triangle', [[439.99, 1234.567]], [[0.0, 0.3]]), 4800, 2),
    Case('triangle_edges', lambda ns: osc(ns, 'Triangle', [[440.0, 6000.0, 12000.0]]), 48000, 3,
         note='6 kHz @ 48 kHz hits the trough exactly -> -0.0 (SURVEY a14)'),
    Case('osc_negative_phase', lambda ns: osc(ns, 'Sawtooth', [[100.0, 0.0]], [[-0.75, -1e-20]]), 2000, 2,
         note='np.mod of negative cycles: fmod+1 branch, incl. the rounds-to-1.0 quirk'),
    Case('gain', lambda ns: gain(ns, osc(ns, 'Triangle', [[440.0, 550.0]]), [[0.2, -0.7]]), 4800, 2),
    Case('mix', _mix, 4800, 3),
    Case('ringmod', _ringmod, 4800, 2),
    Case('amp_int', lambda ns: _amp(ns, [[2.0, 3.0]]), 2400, 2),
    Case('amp_frac', lambda ns: _amp(ns, [[0.5, 1.5]]), 2400, 2, note='negative input ** fractional exp = NaN'),
    Case('broadcast', _broadcast, 2400, 4),
    Case('disabled_osc', _disabled_osc, 512, 2),
    Case('disabled_param', _disabled_param, 512, 2),
    Case('unconnected', _unconnected, 512, 2),
    Case('lowpass_c2_8v', _c2, 48000, 8, tol=1e-4, note='config C2 shape, 8 voices x 1 s'),
    Case('highpass_c2_8v', lambda ns: _c2(ns, 8, 22, 'Sawtooth', 'HighPass'), 48000, 8, tol=1e-4),
    Case('lowpass_test_sigs', _lowpass_test_sigs, 48000, 2, tol=1e-4, note='src/signals/lowpass_test.sigs'),
    Case('lowpass_blockwise', lambda ns: _c2(ns, 4, 7), 512, 4, position=4800, tol=1e-4,
         note='position>0: zero-state restart + 100-frame context warm-up (chain/fx.py:93-105)'),
    Case('lowpass_blockwise_stream', lambda ns: _c2(ns, 4, 7), 3072, 4, position=4800, block=512, tol=1e-4,
         options={'blockwise_reference': 1},
         note='six consecutive 512-frame requests as the reference itself renders them: every request restarts the filter from '
              'zero state 100 frames early (fx.py:82-83, 93-105).  The plan reproduces that with blockwise_reference=1; its '
              'default carries the true state instead (SURVEY 8c: the single-request render is the oracle)'),
    Case('cascade2_seek', _cascade2, 512, 2, position=4800, tol=2e-4,
         note='two chained LowPass nodes, request at position > 0: the reference warms the downstream filter on an upstream '
              'context block that was itself restarted 200 frames early, and the main block on one restarted 100 frames early '
              '(nested restarts); the plan warms the whole chain once from position - 200 with carried state.  Both are '
              'approximations of the stream from position 0; they differ by the upstream transient (8.3e-5 here)'),
    Case('lowpass_order4', lambda ns: lowpass(ns, osc(ns, 'Square', [[220.5, 331.0]]), [[900.0, 2500.0]], order=4),
         24000, 2, tol=1e-4, note='CritFilter.order class knob, chain/fx.py:66'),
    Case('highpass_order3', lambda ns: lowpass(ns, osc(ns, 'Sawtooth', [[220.5, 331.0]]), [[900.0, 2500.0]], 'HighPass', 3),
         24000, 2, tol=1e-4, note='odd order: one first-order section'),
    Case('cascade8', _cascade8, 48000, 4, tol=1e-4, note='config C4 shape: 8 chained LowPass nodes'),
    Case('fanout', _fanout, 4800, 4, tol=1e-4),
    Case('lfo_gain', _lfo_gain, 2400, 2, position=12345, note='Gain.right driven by an oscillator (block-rate port, fx.py:52)'),
    Case('lfo_hertz', _lfo_hertz, 2400, 2, position=3 * RATE + 7,
         note='Osc.hertz / Osc.phase driven by emitters (osc.py:28-30); phase ~ 100 cycles at this position'),
    Case('lfo_mix_amp', _lfo_mix_amp, 2400, 2, position=777, note='Mix.mix and Amp.right modulated (fx.py:39, 59)'),
    Case('lfo_chain', _lfo_chain, 4800, 2, tol=1e-4, note='modulated Gain feeding a LowPass'),
    Case('lfo_blockwise', _lfo_gain, 4096, 2, position=1000, block=512,
         note='8 requests of 512 frames: the LFO is re-sampled at the first frame of each'),
    Case('lfo_hertz_blockwise', _lfo_hertz, 2048, 2, position=9000, block=256,
         note='per-request frequency: the phase jumps between requests exactly as in the reference'),
    Case('lfo_cutoff', _lfo_cutoff, 4800, 2, position=12345, tol=1e-4,
         note='LowPass.cutoff driven by emitters: designed per request at the request position (fx.py:98-102, 124-129); '
              'position > 0, so the 100-frame context warm-up runs with the same design'),
    Case('lfo_cutoff_hp3', _lfo_cutoff_hp3, 4800, 2, position=0, tol=1e-4, note='modulated cutoff, odd order, high-pass'),
    Case('lfo_cutoff_cascade', _lfo_cutoff_cascade, 9600, 2, position=0, tol=1e-4,
         note='four chained filters, every cutoff modulated'),
    Case('lowpass_60s', lambda ns: _c2(ns, 2, 60), 60 * RATE, 2, tol=1e-4, stride=1009,
         note='cascaded-IIR-over-60-s budget; golden keeps every 1009th frame'),
    Case('cascade8_60s', lambda ns: _cascade8(ns, 2, 61), 60 * RATE, 2, tol=1e-4, stride=1009),
]

CASES_BY_NAME = {c.name: c for c in CASES}


# graphs the reference rejects; (name, build, frames, channels, exception type name)
ERROR_CASES = [
    ('scalar_cutoff_multichannel',
     lambda ns: lowpass(ns, osc(ns, 'Sine', [[100.0, 200.0]]), [[500.0]]), 256, 2, 'IndexError'),
    ('cutoff_at_nyquist',
     lambda ns: lowpass(ns, osc(ns, 'Sine', [[100.0]]), [[24000.0]]), 256, 1, 'ValueError'),
    ('cutoff_zero',
     lambda ns: lowpass(ns, osc(ns, 'Sine', [[100.0]]), [[0.0]]), 256, 1, 'ValueError'),
    ('mono_input_multichannel_filter',
     lambda ns: lowpass(ns, osc(ns, 'Sine', [[100.0]]), [[500.0, 600.0]]), 256, 2, 'IndexError'),
]

# graphs that compile but whose RENDER the reference rejects (the parameter only exists at run time)
RUNTIME_ERROR_CASES = [
    ('modulated_cutoff_below_zero',      # an LFO drives the cutoff to -1000 Hz at position 0: butter() rejects Wn = 0 (after the clip)
     lambda ns: _with_cutoff(ns, osc(ns, 'Sine', [[440.0]]), gain(ns, _lfo(ns, 'Sine', [[0.25]], [[0.75]]), [[1000.0]])), 256, 1, 'ValueError'),
]
