"""Graph -> ``sigb_node`` records: the host-side topological sort and shape inference.

Replaces the reference's per-block recursion (BoundPort.request -> Emitter.respond -> _eval,
/root/reference/src/signals/chain/__init__.py:287-300, 253-258) by ONE walk at plan-compile time.
The walk follows the same edges in the same order as the recursion and enforces the same
contracts, so a graph the reference would reject is rejected here with the same exception type:

* every response must be broadcast-compatible with the request (``BadShape``, :292-293);
* the request's channel count flows down unchanged through ``forward`` (:302-303) and is
  re-sliced only by ``Merge`` (chain/shape.py:73-74);
* filters index cutoff and input per requested channel without broadcasting (chain/fx.py:99,105);
* a disabled emitter yields ``zeros((1,1))`` or, for pass-through nodes, its input (:253-254, 416-417);
* an unconnected port yields ``zeros((1,1))`` (:296-298).

The walker is duck-typed on the reference's public surface (class name, ``inputs_by_port``,
``get_state()``, ``channels``), so it lowers ``signals_b200.chain`` nodes and the reference's own
``signals.chain`` node objects alike.
"""
from __future__ import annotations

import dataclasses
import typing

import numpy as np

from signals_b200 import _lib
from signals_b200 import chain as _chain
from signals_b200.chain import BadShape, ChainLayerError, FilterIndexError, UnsupportedGraph

_OSC = {'Sine': _lib.WAVE_SINE, 'Square': _lib.WAVE_SQUARE, 'Sawtooth': _lib.WAVE_SAWTOOTH,
        'Triangle': _lib.WAVE_TRIANGLE}
_FILTER = {'LowPass': _lib.FILT_LOWPASS, 'HighPass': _lib.FILT_HIGHPASS}
_TAPS = ('Wave', 'Spec', 'FileWriter')     # pass-through side-effect nodes (chain/vis.py:61-64, chain/files.py:89-102)
_KNOWN = ('Fixed', *_OSC, 'Mix', 'RingMod', 'Gain', 'Amp', *_FILTER, 'BandPass', 'BandStop', 'Merge',
          'GroupSum', 'PanSum', 'Buffer', *_TAPS)


def node_kind(node) -> str:
    """First known class name in the MRO, so subclasses (e.g. an order-4 LowPass) keep their kind."""
    for klass in type(node).__mro__:
        if klass.__name__ in _KNOWN:
            return klass.__name__
    return type(node).__name__


@dataclasses.dataclass
class GraphRecords:
    nodes: list          # list[_lib.SigbNode]
    data: np.ndarray     # float64 table storage the records index into
    root: int
    channels: int
    rate: int
    buffers: dict        # record index -> Buffer node (bound to device memory by the engine)
    sources: list        # record index -> originating node object (diagnostics)
    # pass-through side-effect nodes: (tap node, requested channels, tap index in the plan | None when the tap sits at
    # the root, where the block it sees IS the rendered output)
    taps: list = dataclasses.field(default_factory=list)
    tracked: bool = True  # every node is a signals_b200.chain node (graph edits bump signals_b200.chain.graph_epoch)
    fixed_values: list = dataclasses.field(default_factory=list)   # (Fixed state object, array, snapshot copy)

    def node_array(self):
        arr = (_lib.SigbNode * len(self.nodes))()
        for i, n in enumerate(self.nodes):
            arr[i] = n
        return arr


class _Lowering:

    def __init__(self, channels: int, rate: int, frames: int):
        self.channels = channels
        self.rate = rate
        self.frames = frames
        self.nodes: list = []
        self.sources: list = []
        self.tables: list[np.ndarray] = []
        self.n_data = 0
        self.buffers: dict = {}
        self.memo: dict = {}
        self.active: set = set()
        self.taps: list = []
        self.n_tap_records = 0
        self.tracked = True
        self.fixed_values: list = []

    # -- record helpers ---------------------------------------------------------------------
    def emit(self, source, kind: int, channels: int, inputs=(-1, -1, -1), subtype: int = 0, order: int = 0,
             context: int = 0, rows: int = 0, data_off: int = 0) -> int:
        rec = _lib.SigbNode()
        rec.kind, rec.subtype, rec.channels = kind, subtype, channels
        padded = tuple(inputs) + (-1,) * (3 - len(inputs))
        for k in range(3):
            rec.inputs[k] = padded[k]
        rec.order, rec.context, rec.rows, rec.data_off = order, context, rows, data_off
        self.nodes.append(rec)
        self.sources.append(source)
        return len(self.nodes) - 1

    def port(self, node, name: str, creq: int, frames: int) -> tuple[int, int]:
        """BoundPort.request at compile time: (record index | -1, channels of the response)."""
        src = node.inputs_by_port.get(name)
        if src is None:
            return -1, 1
        idx, ch = self.visit(src, creq)
        if ch not in (1, creq):
            raise BadShape(src, (frames, ch), (frames, creq))
        return idx, ch

    @staticmethod
    def broadcast(node, *chs: int) -> int:
        wide = {c for c in chs if c != 1}
        if len(wide) > 1:
            raise ValueError(f'operands could not be broadcast together in {node_kind(node)}: channel counts {sorted(wide)}')
        return wide.pop() if wide else 1

    # -- the walk ---------------------------------------------------------------------------
    def visit(self, node, creq: int, at_root: bool = False) -> tuple[int, int]:
        key = (id(node), creq)
        if key in self.memo:
            return self.memo[key]
        if id(node) in self.active:
            raise ChainLayerError('Cycle detected')
        self.active.add(id(node))
        if not isinstance(node, _chain.Signal):
            self.tracked = False          # e.g. the reference's own node objects: no edit hooks, the engine re-walks
        try:
            result = self._lower(node, creq, at_root)
        finally:
            self.active.discard(id(node))
        self.memo[key] = result
        return result

    def _lower(self, node, creq: int, at_root: bool = False) -> tuple[int, int]:
        kind = node_kind(node)
        st = node.get_state()
        if not getattr(st, 'enabled', True):
            flags = type(node).flags()
            passthru = getattr(type(flags), 'PASSTHRU', None)
            if passthru is not None and (flags & passthru) and 'input' in getattr(node, '_ports', {}):
                return self.port(node, 'input', creq, self.frames)
            return self.emit(node, _lib.NODE_ZERO, 1), 1
        F = self.frames
        if kind in _TAPS:
            # the tap's audio result IS its input (PassThroughResult.forward, chain/__init__.py:409-417);
            # queueing blocks for the GUI / writing the file is host-side work outside the render
            src = node.inputs_by_port.get('input')
            if at_root or src is None:
                # at the root the tap sees the rendered block itself; nothing to keep aside
                if not any(t[0] is node for t in self.taps):
                    self.taps.append((node, creq, None))
                if src is None:
                    return -1, 1
                idx, ch = self.visit(src, creq, at_root=True)
                if ch not in (1, creq):
                    raise BadShape(src, (F, ch), (F, creq))
                return idx, ch
            idx, ch = self.port(node, 'input', creq, F)
            self.taps.append((node, creq, self.n_tap_records))
            self.n_tap_records += 1
            return self.emit(node, _lib.NODE_TAP, ch, (idx,)), ch
        if kind == 'Fixed':
            value = np.asarray(st.value, dtype=np.float64)
            rows, ch = value.shape
            if rows != 1:
                raise UnsupportedGraph(f'Fixed with {rows} rows: frame-rate tables are served by signals_b200.chain.ext.Buffer')
            off = self.n_data
            self.fixed_values.append((st, st.value, np.array(st.value, copy=True)))
            self.tables.append(np.ascontiguousarray(value[0]))
            self.n_data += ch
            return self.emit(node, _lib.NODE_FIXED, ch, rows=1, data_off=off), ch
        if kind in _OSC:
            ph, phc = self.port(node, 'phase', creq, 1)     # osc.py:28 requests phase first
            hz, hzc = self.port(node, 'hertz', creq, 1)
            ch = self.broadcast(node, hzc, phc)
            return self.emit(node, _lib.NODE_OSC, ch, (hz, ph), subtype=_OSC[kind]), ch
        if kind == 'Mix':
            m, mc = self.port(node, 'mix', creq, 1)
            le, lc = self.port(node, 'left', creq, F)
            ri, rc = self.port(node, 'right', creq, F)
            ch = self.broadcast(node, mc, lc, rc)
            return self.emit(node, _lib.NODE_MIX, ch, (le, ri, m)), ch
        if kind == 'RingMod':
            le, lc = self.port(node, 'left', creq, F)
            ri, rc = self.port(node, 'right', creq, F)
            ch = self.broadcast(node, lc, rc)
            return self.emit(node, _lib.NODE_RINGMOD, ch, (le, ri)), ch
        if kind in ('Gain', 'Amp'):
            le, lc = self.port(node, 'left', creq, F)
            ri, rc = self.port(node, 'right', creq, 1)
            ch = self.broadcast(node, lc, rc)
            return self.emit(node, _lib.NODE_GAIN if kind == 'Gain' else _lib.NODE_AMP, ch, (le, ri)), ch
        if kind in _FILTER:
            cut, cutc = self.port(node, 'cutoff', creq, 1)
            inp, inc = self.port(node, 'input', creq, F)
            if cutc < creq:
                raise FilterIndexError(f'index {cutc} is out of bounds for axis 1 with size {cutc} '
                                       f'(cutoff of {node.cls_name()!r} is not broadcast over {creq} channels, fx.py:99)')
            if inc < creq:
                raise FilterIndexError(f'index {inc} is out of bounds for axis 1 with size {inc} '
                                       f'(input of {node.cls_name()!r} is not broadcast over {creq} channels, fx.py:105)')
            return self.emit(node, _lib.NODE_FILTER, creq, (inp, cut), subtype=_FILTER[kind],
                             order=int(node.order), context=int(node.context_frames())), creq
        if kind in ('BandPass', 'BandStop'):
            # the reference dies unpacking a scalar at fx.py:99; keep the exception type
            raise TypeError('Value after * must be an iterable, not numpy.float64 (band filters are broken in the '
                            'reference, chain/fx.py:99, and have no defined result)')
        if kind == 'Merge':
            sides = []
            for name in ('left', 'right'):
                src = node.inputs_by_port.get(name)
                if src is None:
                    raise BadShape(node, (F, 1), (F, creq))     # shape.py:69-72 FIXME in the reference
                sc = int(src.channels)
                idx, ch = self.visit(src, sc)
                if ch not in (1, sc):
                    raise BadShape(src, (F, ch), (F, sc))
                sides.append((idx, sc))
            total = sides[0][1] + sides[1][1]
            return self.emit(node, _lib.NODE_MERGE, total, (sides[0][0], sides[1][0])), total
        if kind in ('GroupSum', 'PanSum'):
            src = node.inputs_by_port.get('input')
            if src is None:
                raise BadShape(node, (F, 1), (F, creq))
            sc = int(src.channels)
            idx, ch = self.visit(src, sc)
            if ch not in (1, sc):
                raise BadShape(src, (F, ch), (F, sc))
            if kind == 'GroupSum':
                groups = int(st.groups)
                if sc % groups:
                    raise BadShape(node, (F, sc), (F, groups))
                return self.emit(node, _lib.NODE_GROUPSUM, groups, (idx,), order=groups), groups
            pan, _ = self.port(node, 'pan', sc, 1)
            return self.emit(node, _lib.NODE_PANSUM, 2, (idx, pan)), 2
        if kind == 'Buffer':
            if node.samples is None:
                raise UnsupportedGraph('Buffer without samples')
            rows, ch = int(node.samples.shape[0]), int(node.samples.shape[1])
            idx = self.emit(node, _lib.NODE_BUFFER, ch, rows=min(rows, 2 ** 31 - 1))
            self.buffers[idx] = node
            return idx, ch
        raise UnsupportedGraph(f'{type(node).__module__}.{type(node).__qualname__} has no B200 lowering '
                               f'(supported: {", ".join(_KNOWN)})')


def lower(emitter, channels: int, rate: int, frames: int = 0) -> GraphRecords:
    """Lower the sub-graph under ``emitter`` for a ``(frames, channels)`` request at ``rate`` Hz."""
    lw = _Lowering(int(channels), int(rate), int(frames))
    root, ch = lw.visit(emitter, lw.channels, at_root=True)
    if ch not in (1, lw.channels):
        raise BadShape(emitter, (frames, ch), (frames, channels))
    data = np.concatenate(lw.tables) if lw.tables else np.zeros(0)
    return GraphRecords(nodes=lw.nodes, data=np.ascontiguousarray(data, dtype=np.float64), root=root,
                        channels=lw.channels, rate=lw.rate, buffers=lw.buffers, sources=lw.sources, taps=lw.taps,
                        tracked=lw.tracked, fixed_values=lw.fixed_values)


def signature(emitter) -> tuple:
    """Cheap fingerprint of everything a plan depends on (topology, enable flags, parameter values);
    the engine recompiles when it changes -- the reference re-reads node state on every block."""
    seen: dict = {}
    out: list = []

    def walk(node):
        if id(node) in seen:
            return
        seen[id(node)] = True
        st = node.get_state()
        items: list = [id(node), type(node).__qualname__, getattr(st, 'enabled', True)]
        kind = node_kind(node)
        if kind == 'Fixed':
            v = st.value
            items += [v.shape, v.ctypes.data,
                      v.tobytes() if v.size <= 4096 else (v.ravel()[:: max(1, v.size // 1024)].tobytes(), float(v.sum()))]
        elif kind == 'GroupSum':
            items.append(st.groups)
        elif kind == 'Buffer':
            s = node.samples
            items += [id(s), tuple(s.shape) if s is not None else None]
        elif kind in _FILTER:
            items += [node.order, node.context_frames()]
        ports = getattr(node, 'inputs_by_port', {})
        for name in sorted(ports):
            items.append((name, id(ports[name])))
        out.append(tuple(items))
        for name in sorted(ports):
            walk(ports[name])

    walk(emitter)
    return tuple(out)
