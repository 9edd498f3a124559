"""``.sigs`` patch files -> node graphs, headless.

The reference saves a patch as the list of commands that rebuilds it (map/control.py:561-594, 807-823) and
replays them through its ``Controller``/``Map`` (grid, undo stack, Qt).  This loader executes the
graph-building subset of that command language directly against ``signals_b200.chain`` nodes:

    sink   <coord> <device>                map/control.py:493-511
    source <coord> <device>                map/control.py:145-173   (kept as an unconnected placeholder)
    +      <coord> <qualname> [k=v ...]    map/control.py:291-330
    *      <coord> k=v [k=v ...]           map/control.py:380-420   (edit state)
    >      <src> <dst>.<port>              map/control.py:422-452
    >/     <dst>.<port>                    map/control.py:454-480   (disconnect)
    -      <coord>                         map/control.py:332-378

Values use the reference's syntax (map/__init__.py:104-148): JSON where it parses (lists become numpy
arrays), the raw string otherwise.  Coordinates are ``<row><column-letters>`` (map/__init__.py:55-101); they
are opaque keys here.  Blank lines and ``#`` comments are skipped.
"""
from __future__ import annotations

import json
import re
import shlex
import typing

import attr
import numpy as np

from signals_b200 import SignalsError
from signals_b200.chain import BlockLoc, Emitter, Receiver, Shape, Signal
from signals_b200.chain import dev as dev_mod
from signals_b200.chain import discovery

_COORD = re.compile(r'^\d+[a-z]+$')


class PatchError(SignalsError):
    pass


def parse_value(text: str):
    """SigStateItem.parse_value, map/__init__.py:131-140."""
    try:
        v = json.loads(text)
    except ValueError:
        return text
    return np.array(v) if isinstance(v, list) else v


def _apply_state(sig: Signal, items: typing.Iterable[str]) -> None:
    updates = {}
    for item in items:
        if '=' not in item:
            raise PatchError(f'bad state item {item!r}')
        k, v = item.split('=', 1)
        updates[k] = parse_value(v)
    if updates:
        try:
            sig.set_state(attr.evolve(sig.get_state(), **updates))
        except TypeError as e:
            raise PatchError(f'{sig.cls_name()}: {e}')


class Patch:
    """The nodes of a patch by grid coordinate, and its sinks."""

    def __init__(self):
        self.nodes: dict[str, Signal] = {}
        self.sinks: dict[str, dev_mod.SinkDevice] = {}
        self.rack = discovery.Rack()
        self.rack.scan()

    # -- commands -----------------------------------------------------------------------------
    def _coord(self, text: str) -> str:
        if not _COORD.match(text):
            raise PatchError(f'bad coordinates {text!r}')
        return text

    def _at(self, coord: str) -> Signal:
        try:
            return self.nodes[coord]
        except KeyError:
            raise PatchError(f'nothing at {coord}')

    def execute(self, line: str) -> None:
        tokens = shlex.split(line, comments=True)
        if not tokens:
            return
        cmd, args = tokens[0], tokens[1:]
        if cmd == 'sink':
            at = self._coord(args[0])
            sink = dev_mod.SinkDevice(self.rack.get_sink(args[1]))
            _apply_state(sink, args[2:])
            self.nodes[at] = self.sinks[at] = sink
        elif cmd == 'source':
            raise PatchError('source devices (microphones) have no B200 lowering')
        elif cmd == '+':
            at = self._coord(args[0])
            if at in self.nodes:
                raise PatchError(f'{at} is occupied')
            sig = discovery.load_signal(args[1])()
            _apply_state(sig, args[2:])
            self.nodes[at] = sig
        elif cmd == '*':
            _apply_state(self._at(self._coord(args[0])), args[1:])
        elif cmd == '>':
            src = self._at(self._coord(args[0]))
            dst_at, _, port_name = args[1].partition('.')
            dst = self._at(self._coord(dst_at))
            if not isinstance(src, Emitter):
                raise PatchError(f'{args[0]} does not emit')
            if not isinstance(dst, Receiver) or port_name not in dst.port_names():
                raise PatchError(f'{args[1]}: no such port')
            setattr(dst, port_name, src)
        elif cmd == '>/':
            dst_at, _, port_name = args[0].partition('.')
            delattr(self._at(self._coord(dst_at)), port_name)
        elif cmd == '-':
            at = self._coord(args[0])
            self._at(at).destroy()
            del self.nodes[at]
            self.sinks.pop(at, None)
        else:
            raise PatchError(f'command {cmd!r} does not build graphs (only sink source + * > >/ - are replayed)')

    # -- rendering ----------------------------------------------------------------------------
    def root(self, sink_at: typing.Optional[str] = None) -> tuple[dev_mod.SinkDevice, Emitter]:
        if sink_at is None:
            if len(self.sinks) != 1:
                raise PatchError(f'patch has {len(self.sinks)} sinks; name one')
            sink_at = next(iter(self.sinks))
        sink = self.sinks[sink_at]
        emitter = sink.inputs_by_port.get('input')
        if emitter is None:
            raise PatchError(f'sink {sink_at} has no input')
        return sink, emitter

    def render(self, position: int, frames: int, rate: int = 48000, sink_at: typing.Optional[str] = None,
               taps: bool = False) -> np.ndarray:
        """What the sink's callback would deliver for ``frames`` frames from ``position``: float32
        ``(frames, sink channels)``, rendered by the CUDA path."""
        from signals_b200 import engine
        sink, emitter = self.root(sink_at)
        loc = BlockLoc(position=position, rate=rate, shape=Shape(frames=frames, channels=sink.get_state().channels))
        block = engine.default_engine().render(emitter, loc)
        if taps:        # FileWriter / Wave nodes on the way get their blocks too (chain/files.py:99-102, chain/vis.py:61-64)
            engine.default_engine().serve_taps(emitter, loc, rendered=np.broadcast_to(block, tuple(loc.shape)))
        return np.broadcast_to(block, tuple(loc.shape)).astype(np.float32)


def loads(text: str) -> Patch:
    patch = Patch()
    for n, line in enumerate(text.splitlines(), 1):
        try:
            patch.execute(line)
        except (IndexError, SignalsError) as e:
            raise PatchError(f'line {n}: {line.strip()!r}: {e}') from e
    return patch


def load(path) -> Patch:
    with open(path) as f:
        return loads(f.read())
