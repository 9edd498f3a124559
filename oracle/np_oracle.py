"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's block-render path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product
(``signals_b200``) never does; it fails loudly without its CUDA library.

This is a numpy restatement of noah-aviel-dove/signals' per-block DSP, op for
op and in the reference's float64 op order.  Citations are
``/root/reference/src/signals/...`` file:line.

Third-party arithmetic the reference delegates to (absent from /root/reference,
pinned in its requirements.txt:5-6): ``numpy==1.23.0`` (sin, mod, sign,
copysign, hstack) and ``scipy==1.10.1`` (``scipy.signal.butter``,
``scipy.signal.sosfilt``).  This image has numpy 2.3.5 / scipy 1.18.1; the
functions used are stable across those versions to the last ulp or so, far
inside the 1e-6 / 1e-4 parity tolerances.  ``butter2_closed_form`` and
``sosfilt_df2t`` below restate the two scipy algorithms (bilinear-transformed
2nd-order Butterworth; cascaded direct-form-II-transposed biquads) and are
checked against scipy in tests/test_oracle.py.

PARITY PINNING: the reference has no golden vectors or known-answer tests for
this path.  The oracle is therefore pinned against
outputs of the reference itself, executed in the build container through
``oracle/ref_harness.py``; those outputs are committed under ``tests/golden/``
together with the generator ``oracle/make_golden.py``.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.signal

# ----------------------------------------------------------------------------
# node-level restatements
# ----------------------------------------------------------------------------


def frame_range(position: int, frames: int) -> np.ndarray:
    """BlockLoc.frame_range, chain/__init__.py:121-125: (F,1) int64 absolute indices."""
    return np.arange(position, position + frames).reshape(-1, 1)


def osc_cycles(position: int, frames: int, rate: int, hertz: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """Osc._eval, chain/osc.py:32: ``frame_range / rate * hertz + phase`` in that op order."""
    return frame_range(position, frames) / rate * hertz + phase


def sine(t: np.ndarray) -> np.ndarray:
    """Sine._osc, chain/osc.py:43."""
    return np.sin(t * 2 * np.pi)


def square(t: np.ndarray) -> np.ndarray:
    """Square._osc, chain/osc.py:49."""
    return np.sign(0.5 - np.mod(t, 1))


def sawtooth(t: np.ndarray) -> np.ndarray:
    """Sawtooth._osc, chain/osc.py:55."""
    return 2 * np.mod(t - 0.5, 1) - 1


def triangle(t: np.ndarray) -> np.ndarray:
    """Triangle._osc, chain/osc.py:61-62 (including the -0.0 at the exact trough)."""
    t = t - 0.25
    return (4 * np.mod(t, 0.5) - 1) * np.sign(np.mod(t, 1) - 0.5)


WAVEFORMS = {'Sine': sine, 'Square': square, 'Sawtooth': sawtooth, 'Triangle': triangle}


def mix(mix_, left, right):
    """Mix._eval, chain/fx.py:39-40."""
    return mix_ * left + (1 - mix_) * right


def ringmod(left, right):
    """RingMod._eval, chain/fx.py:46."""
    return left * right


def gain(left, right):
    """Gain._eval, chain/fx.py:52."""
    return left * right


def amp(input_, exp):
    """Amp._eval, chain/fx.py:58-60."""
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.copysign(input_ ** exp, input_)


FILTER_TYPES = {'LowPass': 'lp', 'HighPass': 'hp', 'BandPass': 'bp', 'BandStop': 'bs'}


def butter_sos(btype: str, order: int, wn) -> np.ndarray:
    """CritFilter._get_sos, chain/fx.py:115-121 (delegates to scipy.signal.butter)."""
    return scipy.signal.butter(N=order, Wn=wn, btype=btype, output='sos')


def butter2_closed_form(btype: str, wn: float) -> np.ndarray:
    """Published algorithm behind ``butter(2, wn, btype, output='sos')``: analog
    prototype 1/(s^2+sqrt(2)s+1), pre-warped K=tan(pi*wn/2), bilinear transform."""
    k = math.tan(math.pi * wn / 2.0)
    n = 1.0 / (1.0 + math.sqrt(2.0) * k + k * k)
    a1 = 2.0 * (k * k - 1.0) * n
    a2 = (1.0 - math.sqrt(2.0) * k + k * k) * n
    if btype == 'lp':
        b = (k * k * n, 2.0 * k * k * n, k * k * n)
    elif btype == 'hp':
        b = (n, -2.0 * n, n)
    else:
        raise ValueError(btype)
    return np.array([[b[0], b[1], b[2], 1.0, a1, a2]])


def sosfilt_df2t(sos: np.ndarray, x: np.ndarray, zi: np.ndarray | None = None):
    """Published algorithm of scipy.signal.sosfilt: cascade of direct-form-II-transposed
    biquads, ``y=b0 x+z0; z0=b1 x-a1 y+z1; z1=b2 x-a2 y`` per section, zero initial state.
    ``x`` is 1-D (one channel) or (F,C) with per-column independent filtering."""
    x = np.asarray(x, dtype=np.float64)
    squeeze = x.ndim == 1
    y = x.reshape(len(x), -1).copy()
    n_sec = sos.shape[0]
    z = np.zeros((n_sec, 2, y.shape[1])) if zi is None else np.array(zi, dtype=np.float64)
    for n in range(y.shape[0]):
        v = y[n]
        for s in range(n_sec):
            b0, b1, b2, _, a1, a2 = sos[s]
            o = b0 * v + z[s, 0]
            z[s, 0] = b1 * v - a1 * o + z[s, 1]
            z[s, 1] = b2 * v - a2 * o
            v = o
        y[n] = v
    return (y[:, 0] if squeeze else y), z


def crit_filter(ctx: np.ndarray, crit_1: np.ndarray, rate: int, btype: str, order: int,
                frames: int, context_frames: int, crit_2: np.ndarray | None = None,
                channels: int | None = None) -> np.ndarray:
    """CritFilter._filter, chain/fx.py:85-106.

    ``ctx`` = [before(context)?, block, after(context)] concatenated along frames
    (chain/__init__.py:308-315); per channel: Wn = clip(crit/(rate/2), 0, 1), butter(),
    sosfilt from zero state, keep ``[-(frames+context):-context]``.
    Indexing is un-broadcast like the reference's (``crit_1[0, i]``, ``ctx[:, i]``).
    """
    if channels is None:
        channels = ctx.shape[1]
    result = np.empty((frames, channels))     # request.loc.shape, fx.py:95-96
    for i in range(channels):
        crits = (crit_1[0, i],) if crit_2 is None else (crit_1[0, i], crit_2[0, i])
        scaled = np.array(crits, dtype=float)
        scaled /= rate / 2
        scaled.clip(0, 1, out=scaled)
        sos = butter_sos(btype, order, scaled)
        sl = slice(-(frames + context_frames), -context_frames)
        result[:, i] = scipy.signal.sosfilt(sos, ctx[:, i], axis=0)[sl]
    return result


def merge(left: np.ndarray, right: np.ndarray) -> np.ndarray:
    """Merge._eval, chain/shape.py:73-74."""
    return np.hstack((left, right))


def group_sum(x: np.ndarray, groups: int) -> np.ndarray:
    """Harness-defined N->G mixdown (the reference's Flatten is broken):
    channel c belongs to group c // (C/groups); float64 sum."""
    f, c = x.shape
    return x.reshape(f, groups, c // groups).sum(-1)


def pan_sum(x: np.ndarray, pan: np.ndarray) -> np.ndarray:
    """Harness-defined stereo mixdown: L = sum((1-pan) y), R = sum(pan y) in float64."""
    return np.stack(((x * (1 - pan)).sum(-1), (x * pan).sum(-1)), axis=-1)


# ----------------------------------------------------------------------------
# graph-level: the reference's pull recursion, restated over duck-typed nodes
# ----------------------------------------------------------------------------

_KNOWN = ('Fixed', 'Sine', 'Square', 'Sawtooth', 'Triangle', 'Mix', 'RingMod', 'Gain', 'Amp',
          'LowPass', 'HighPass', 'BandPass', 'BandStop', 'Merge', 'GroupSum', 'PanSum', 'Buffer')


def _cls(node) -> str:
    """Node kind = first known class name in the MRO (subclasses such as an order-4 LowPass keep their kind)."""
    for klass in type(node).__mro__:
        if klass.__name__ in _KNOWN:
            return klass.__name__
    return type(node).__name__


def _zeros():
    """Emitter.empty_result, chain/__init__.py:249-251."""
    return np.zeros((1, 1))


def _check_shape(block: np.ndarray, frames: int, channels: int, node):
    """BoundPort._do_request shape contract, chain/__init__.py:292-293 / Shape.__ge__ :62-63."""
    if block.ndim != 2 or block.shape[0] not in (1, frames) or block.shape[1] not in (1, channels):
        raise ValueError(f'BadShape: {_cls(node)} returned {block.shape} for request ({frames}, {channels})')
    return block


class GraphOracle:
    """Evaluates a graph of nodes exposing the reference's public surface
    (``inputs_by_port``, ``get_state()``, ``channels``) with the reference's
    recursion: BoundPort.request (chain/__init__.py:296-300), Emitter.respond
    (:253-258), forward_at_block_rate (:305-306), forward_with_context (:308-315).
    Works on the reference's own node objects and on ``signals_b200.chain`` mirrors.
    The per-node block cache (chain/__init__.py:424-457) only de-duplicates
    requests and is not restated (rendering is a pure function of the request).
    """

    def __init__(self, rate: int = 48000):
        self.rate = rate

    # BoundPort.request
    def request(self, node, port: str, position: int, frames: int, channels: int) -> np.ndarray:
        src = node.inputs_by_port.get(port)
        if src is None:
            return _zeros()
        return _check_shape(self.respond(src, position, frames, channels), frames, channels, src)

    def at_block_rate(self, node, port, position, frames, channels):
        return self.request(node, port, position, 1, channels)

    def with_context(self, node, port, position, frames, channels, ctx):
        blocks = []
        if position > 0:
            blocks.append(self.request(node, port, max(position - ctx, 0), min(ctx, position), channels))
        blocks.append(self.request(node, port, position, frames, channels))
        blocks.append(self.request(node, port, position + frames, ctx, channels))
        return np.concatenate(blocks)

    def respond(self, node, position: int, frames: int, channels: int) -> np.ndarray:
        name = _cls(node)
        enabled = getattr(node.get_state(), 'enabled', True)
        if not enabled:
            return _zeros()
        if name == 'Fixed':
            return node.get_state().value  # chain/fixed.py:38-39
        if name in WAVEFORMS:
            phase = self.at_block_rate(node, 'phase', position, frames, channels)
            hertz = self.at_block_rate(node, 'hertz', position, frames, channels)
            return WAVEFORMS[name](osc_cycles(position, frames, self.rate, hertz, phase))
        if name == 'Mix':
            m = self.at_block_rate(node, 'mix', position, frames, channels)
            return mix(m, self.request(node, 'left', position, frames, channels),
                       self.request(node, 'right', position, frames, channels))
        if name == 'RingMod':
            return ringmod(self.request(node, 'left', position, frames, channels),
                           self.request(node, 'right', position, frames, channels))
        if name == 'Gain':
            return gain(self.request(node, 'left', position, frames, channels),
                        self.at_block_rate(node, 'right', position, frames, channels))
        if name == 'Amp':
            return amp(self.request(node, 'left', position, frames, channels),
                       self.at_block_rate(node, 'right', position, frames, channels))
        if name in ('LowPass', 'HighPass'):
            crit = self.at_block_rate(node, 'cutoff', position, frames, channels)
            ctxf = node.context_frames()
            ctx = self.with_context(node, 'input', position, frames, channels, ctxf)
            ctx = np.asarray(ctx)
            return crit_filter(ctx, crit, self.rate, FILTER_TYPES[name], node.order, frames, ctxf, channels=channels)
        if name == 'Merge':
            lc = node.inputs_by_port['left'].channels
            rc = node.inputs_by_port['right'].channels
            return merge(self.request(node, 'left', position, frames, lc),
                         self.request(node, 'right', position, frames, rc))
        if name == 'GroupSum':  # signals_b200 extension; harness-defined oracle
            c = node.inputs_by_port['input'].channels
            x = np.broadcast_to(self.request(node, 'input', position, frames, c), (frames, c))
            return group_sum(x, node.get_state().groups)
        if name == 'PanSum':  # signals_b200 extension; harness-defined oracle
            c = node.inputs_by_port['input'].channels
            x = np.broadcast_to(self.request(node, 'input', position, frames, c), (frames, c))
            pan = self.at_block_rate(node, 'pan', position, frames, c)
            return pan_sum(x, np.broadcast_to(pan, (1, c)))
        raise NotImplementedError(f'oracle: unsupported node {name}')

    def render(self, emitter, position: int, frames: int, channels: int) -> np.ndarray:
        """Root pull (chain/dev.py:173) broadcast to the full requested shape (:178)."""
        block = _check_shape(self.respond(emitter, position, frames, channels), frames, channels, emitter)
        return np.array(np.broadcast_to(block, (frames, channels)), dtype=np.float64)


# ----------------------------------------------------------------------------
# array-level renders of the BASELINE configs (used by bench.py cpu_baseline and tests)
# ----------------------------------------------------------------------------

def example_sine_block(start_idx: int, frames: int, samplerate: float, frequency: float = 500.0,
                       amplitude: float = 0.2) -> np.ndarray:
    """scripts/example_sine.py:50-53 callback body (config C1's CPU formula)."""
    t = (start_idx + np.arange(frames)) / samplerate
    t = t.reshape(-1, 1)
    return amplitude * np.sin(2 * np.pi * frequency * t)


def render_voice_chain(position: int, frames: int, rate: int, hertz, phase, cutoff, gain_,
                       wave: str = 'Sine', btype: str | None = 'lp', order: int = 2,
                       gain_before_filter: bool = False) -> np.ndarray:
    """Config C2 as arrays: osc -> (biquad) -> gain, single request from zero state.
    hertz/phase/cutoff/gain_ are (V,) float64.  Equivalent to the GraphOracle on the
    Sine<-Fixed -> LowPass<-Fixed -> Gain<-Fixed graph at position 0."""
    hertz = np.asarray(hertz, dtype=float).reshape(1, -1)
    phase = np.asarray(phase, dtype=float).reshape(1, -1)
    x = WAVEFORMS[wave](osc_cycles(position, frames, rate, hertz, phase))
    g = np.asarray(gain_, dtype=float).reshape(1, -1)
    if gain_before_filter:
        x = gain(x, g)
    if btype is not None:
        y = np.empty_like(x)
        for i, fc in enumerate(np.asarray(cutoff, dtype=float).ravel()):
            wn = min(max(fc / (rate / 2), 0.0), 1.0)
            y[:, i] = scipy.signal.sosfilt(butter_sos(btype, order, wn), x[:, i], axis=0)
        x = y
    if not gain_before_filter:
        x = gain(x, g)
    return x


def render_cascade(x: np.ndarray, cutoffs: np.ndarray, rate: int, btype: str = 'lp', order: int = 2,
                   zi: np.ndarray | None = None):
    """Config C4 as arrays: ``cutoffs`` is (S, C); S chained filter nodes per channel.
    Returns (y, zf) with zf shaped (S, C, n_sections, 2) so a stream can be continued."""
    n_s, c = cutoffs.shape
    y = np.array(x, dtype=np.float64)
    n_sec = (order + 1) // 2
    zf = np.zeros((n_s, c, n_sec, 2))
    for s in range(n_s):
        for i in range(c):
            wn = min(max(cutoffs[s, i] / (rate / 2), 0.0), 1.0)
            sos = butter_sos(btype, order, wn)
            z0 = np.zeros((n_sec, 2)) if zi is None else zi[s, i]
            y[:, i], zf[s, i] = scipy.signal.sosfilt(sos, y[:, i], axis=0, zi=z0)
    return y, zf


def render_bank(position: int, frames: int, rate: int, hertz, phase, amp, groups: int, chunk: int = 256) -> np.ndarray:
    """Config C3 as arrays: group_sum(gain(sine(cycles), amp)) accumulated in float64, partials
    processed `chunk` at a time to bound the (frames, chunk) float64 temporaries."""
    hertz, phase, amp = (np.asarray(a, dtype=float) for a in (hertz, phase, amp))
    p = hertz.size
    per = p // groups
    out = np.zeros((frames, groups))
    for g in range(groups):
        for a in range(g * per, (g + 1) * per, chunk):
            b = min(a + chunk, (g + 1) * per)
            x = gain(sine(osc_cycles(position, frames, rate, hertz[None, a:b], phase[None, a:b])), amp[None, a:b])
            out[:, g] += x.sum(-1)
    return out


def render_instances(prm: dict, position: int, frames: int, rate: int, chunk: int = 256) -> np.ndarray:
    """Config C5 as arrays: per instance osc(wave) -> (LowPass | HighPass | none) -> gain, then
    pan_sum; single request from zero state (position must be 0 when any instance has a filter)."""
    waves = ('Sine', 'Square', 'Sawtooth', 'Triangle')
    out = np.zeros((frames, 2))
    n = prm['hertz'].size
    for a in range(0, n, chunk):
        b = min(a + chunk, n)
        y = np.empty((frames, b - a))
        for j in range(a, b):
            btype = (None, 'lp', 'hp')[int(prm['filt'][j])]
            y[:, j - a] = render_voice_chain(position, frames, rate, prm['hertz'][j:j + 1], prm['phase'][j:j + 1],
                                             prm['cutoff'][j:j + 1], prm['gain'][j:j + 1],
                                             wave=waves[int(prm['wave'][j])], btype=btype)[:, 0]
        out += pan_sum(y, prm['pan'][None, a:b])
    return out
