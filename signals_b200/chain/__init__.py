"""Node-graph core: the host-side mirror of ``signals.chain``.

Same names, argument meaning and error behaviour as the reference's graph core
(/root/reference/src/signals/chain/__init__.py), so graphs are built exactly as before::

    sine = Sine(); sine.hertz = fixed          # chain/__init__.py:367-377 (port descriptors)
    block = sink.input.request(loc)            # chain/__init__.py:296-300 (the root pull)

What differs is *where* a request is evaluated.  The reference recurses node by node and
materialises a float64 numpy block per edge (:253-258, :287-315).  Here a request compiles the
sub-graph under the requested emitter into a fused CUDA launch plan (``signals_b200.plan``) and
renders it on the GPU (``signals_b200.engine``); built-in nodes carry no numpy arithmetic and there
is no CPU fallback.
"""
from __future__ import annotations

import abc
import enum
import typing

import attr
import attrs.validators
import numpy as np

from signals_b200 import PortName, SignalFlags, SignalsError


class ChainLayerError(SignalsError):
    pass


class Shape(typing.NamedTuple):
    """``(frames, channels)``; ``<=`` / ``>=`` mean *broadcast-compatible* (reference :25-84).

    >>> s = Shape(frames=10, channels=2)
    >>> (1, 1) <= Shape(frames=10, channels=1) <= s
    True
    >>> Shape(frames=3, channels=2) <= s
    False
    >>> (10, 1) <= s        # a plain tuple on the left dispatches to Shape.__ge__
    True
    """
    frames: int
    channels: int

    @classmethod
    def unit(cls) -> 'Shape':
        return cls(frames=1, channels=1)

    @classmethod
    def of_array(cls, array: np.ndarray) -> 'Shape':
        return cls(*array.shape)

    @staticmethod
    def _fits(small, big) -> bool:
        return all(s == 1 or s == b for s, b in zip(small, big))

    def __le__(self, other) -> bool:
        return self._fits(self, other)

    def __ge__(self, other) -> bool:
        return self._fits(other, self)


class BadShape(ChainLayerError):

    def __init__(self, source, shape: tuple, constraint: tuple):
        name = source.cls_name() if hasattr(source, 'cls_name') else str(source)
        super().__init__(f'Invalid response from {name!r}): '
                         f'Block with shape {tuple(shape)} incompatible with requested shape {tuple(constraint)}')


class BadStateSchema(ChainLayerError):

    def __init__(self, sig, state):
        super().__init__(f'Signal {sig.cls_name()!r} cannot accept state of type {state.cls_name()!r}')


class BadStateValue(ChainLayerError):

    def __init__(self, state, key: str, value, reason=None):
        reason = '' if reason is None else f': ({reason})'
        super().__init__(f'Value {value!r} is invalid for property {key!r} in schema {state.cls_name()!r}{reason}')


class UnsupportedGraph(ChainLayerError):
    """The B200 evaluator cannot lower this node or graph shape (raised at plan-compile time)."""


class FilterIndexError(ChainLayerError, IndexError):
    """A filter's cutoff or input is narrower than the requested channel count.

    The reference indexes ``crit_1[0, i]`` / ``input_[:, i]`` without broadcasting
    (chain/fx.py:99,105) and dies with IndexError; this is that error."""


class FilterDesignError(ChainLayerError, ValueError):
    """Critical frequency outside (0, Nyquist): scipy.signal.butter's ValueError via chain/fx.py:102."""


@attr.s(auto_attribs=True, frozen=True, kw_only=True, order=False)
class BlockLoc:
    """Absolute address of a block: ``position`` (frames), ``rate`` (Hz), ``shape`` (reference :107-159)."""
    position: int
    rate: int
    shape: Shape

    @property
    def end_position(self) -> int:
        return self.position + self.shape[0]

    @property
    def timestamp(self) -> float:
        return self.position / self.rate

    @property
    def frame_range(self) -> np.ndarray:
        frames = np.arange(self.position, self.end_position).reshape(-1, 1)
        frames.flags.writeable = False
        return frames

    def _with_shape(self, frames: int, channels: int) -> 'BlockLoc':
        if (frames, channels) == tuple(self.shape):
            return self
        return attr.evolve(self, shape=Shape(frames=frames, channels=channels))

    def resize(self, new_frames: int) -> 'BlockLoc':
        return self._with_shape(new_frames, self.shape[1])

    def reslice(self, new_channels: int) -> 'BlockLoc':
        return self._with_shape(self.shape[0], new_channels)

    def __le__(self, other: 'BlockLoc') -> bool:
        """Containment: same rate, frame span inside ``other``, no more channels."""
        return (self.rate == other.rate
                and other.position <= self.position
                and self.end_position <= other.end_position
                and self.shape[1] <= other.shape[1])

    def before(self, frames: int) -> 'BlockLoc':
        start = max(self.position - frames, 0)
        return attr.evolve(self, position=start, shape=Shape(frames=self.position - start, channels=self.shape[1]))

    def after(self, frames: int) -> 'BlockLoc':
        return attr.evolve(self, position=self.end_position, shape=Shape(frames=frames, channels=self.shape[1]))


@attr.s(auto_attribs=True, frozen=True, kw_only=True)
class Request:
    requestor: 'Receiver'
    port: PortName
    loc: BlockLoc


class RequestRate(enum.Enum):
    UNKNOWN = enum.auto()
    BLOCK = enum.auto()
    FRAME = enum.auto()
    UNUSED_FRAME = enum.auto()


# Graph epoch: bumped by every edit the evaluator can observe -- a port (dis)connected, a state attribute assigned, a
# state object replaced.  The engine keeps a compiled plan for as long as the epoch stands still (one integer compare
# per audio callback instead of a walk over the graph); the reference has nothing to invalidate because it re-reads
# every node's state on every block.  In-place writes INTO a Fixed's value array are invisible here; the engine
# compares those arrays against its snapshot (signals_b200.engine.Engine.plan_for).
_epoch = 0


def graph_epoch() -> int:
    return _epoch


def touch_graph() -> None:
    global _epoch
    _epoch += 1


def _on_state_setattr(instance, attribute, value):
    touch_graph()
    return value


state = attr.s(auto_attribs=True, frozen=False, kw_only=True, on_setattr=_on_state_setattr)


class Named:
    """Dotted class name used by the patch language (mirrors signals.discovery.Named)."""

    @classmethod
    def cls_name(cls) -> str:
        return f'{cls.__module__}.{cls.__qualname__}'


class Signal(abc.ABC, Named):
    @state
    class State(Named):
        pass

    def __init__(self):
        self._state = self.State()

    @classmethod
    @abc.abstractmethod
    def flags(cls) -> SignalFlags:
        return SignalFlags(0)

    @classmethod
    def state_attrs(cls) -> typing.AbstractSet[str]:
        return attr.fields_dict(cls.State).keys()

    def get_state(self):
        return self._state

    def set_state(self, new_state) -> None:
        if not isinstance(new_state, self.State):
            raise BadStateSchema(self, new_state)
        self._state = new_state
        touch_graph()

    def destroy(self) -> None:
        pass


class Emitter(Signal, abc.ABC):
    @state
    class State(Signal.State):
        enabled: bool = attr.ib(validator=attrs.validators.instance_of(bool), default=True)

    def __init__(self):
        super().__init__()
        self._outputs: set[tuple[PortName, 'Receiver']] = set()
        self._last_request: typing.Optional[Request] = None

    @property
    def outputs_with_ports(self):
        return self._outputs

    @property
    def rate(self) -> RequestRate:
        if self._last_request is None:
            return RequestRate.UNKNOWN
        frames = self._last_request.loc.shape.frames
        if frames <= 0:
            return RequestRate.UNKNOWN
        return RequestRate.BLOCK if frames == 1 else RequestRate.FRAME

    @property
    @abc.abstractmethod
    def channels(self) -> int:
        raise NotImplementedError

    def _eval(self, request: Request) -> np.ndarray:
        """Render this emitter's sub-graph for ``request.loc`` on the GPU.

        The reference makes this abstract and implements it per node in numpy; here every
        built-in node shares this one implementation because evaluation is a property of the
        compiled graph, not of a node."""
        from signals_b200 import engine
        return engine.default_engine().render(self, request.loc)

    @classmethod
    def empty_result(cls) -> np.ndarray:
        return np.zeros(Shape.unit())

    def _get_result(self, request: Request) -> np.ndarray:
        return self._eval(request) if self._state.enabled else self.empty_result()

    def respond(self, request: Request) -> np.ndarray:
        self._last_request = request
        return self._get_result(request)

    def destroy(self) -> None:
        super().destroy()
        for port_name, receiver in tuple(self._outputs):
            delattr(receiver, port_name)


class Receiver(Signal, abc.ABC):

    class BoundPort:
        """One input socket of a receiver (reference :266-322)."""

        def __init__(self, parent: 'Receiver', name: PortName, emitter: typing.Optional[Emitter] = None):
            self.name = name
            self.parent = parent
            self.sig = emitter

        def __bool__(self) -> bool:
            return self.sig is not None

        def expel(self) -> None:
            self.sig._outputs.discard((self.name, self.parent))
            self.sig = None
            touch_graph()

        def assign(self, input_: Emitter) -> None:
            if self.sig is not None:
                self.expel()
            self.sig = input_
            input_._outputs.add((self.name, self.parent))
            touch_graph()

        def request(self, loc: BlockLoc) -> np.ndarray:
            """The pull: an unconnected port yields ``zeros((1, 1))``; otherwise the emitter's
            block, which must be broadcast-compatible with ``loc.shape`` (reference :287-300)."""
            if self.sig is None:
                return Emitter.empty_result()
            block = self.sig.respond(Request(requestor=self.parent, port=self.name, loc=loc))
            if not (tuple(block.shape) <= loc.shape):
                raise BadShape(self.sig, block.shape, loc.shape)
            return block

        def forward(self, request: Request) -> np.ndarray:
            return self.request(request.loc)

        def forward_at_block_rate(self, request: Request) -> np.ndarray:
            return self.request(request.loc.resize(1))

        def forward_with_context(self, request: Request, context_frames: int) -> np.ndarray:
            loc = request.loc
            parts = [self.request(loc.before(context_frames))] if loc.position > 0 else []
            parts += [self.request(loc), self.request(loc.after(context_frames))]
            return np.concatenate([np.broadcast_to(p, (p.shape[0], loc.shape.channels)) for p in parts])

        @property
        def channels(self) -> typing.Optional[int]:
            return None if self.sig is None else self.sig.channels

    def __init__(self):
        super().__init__()
        self._ports = {name: self.BoundPort(parent=self, name=name) for name in self.port_names()}

    @classmethod
    def port_names(cls) -> list[PortName]:
        return [k for k in dir(cls) if isinstance(getattr(cls, k, None), _Port)]

    @property
    def inputs_by_port(self) -> dict[PortName, Emitter]:
        return {p.name: p.sig for p in self._ports.values() if p}

    def upstream(self) -> typing.Sequence[Emitter]:
        """Post-order list of upstream *receivers* ending with ``self`` (reference :347-358)."""
        order: list = []
        self._walk_upstream(set(), set(), order)
        return order

    def _walk_upstream(self, done: set, active: set, order: list) -> None:
        assert id(self) not in active, 'Cycle detected'
        active.add(id(self))
        for inp in self.inputs_by_port.values():
            if isinstance(inp, Receiver) and id(inp) not in done:
                inp._walk_upstream(done, active, order)
        active.discard(id(self))
        done.add(id(self))
        order.append(self)

    def destroy(self) -> None:
        super().destroy()
        for port_name, bound in tuple(self._ports.items()):
            if bound:
                delattr(self, port_name)


class _Port(property):
    pass


def port(name: PortName) -> _Port:
    """Declare an input port: ``getattr`` -> BoundPort, ``setattr`` connects, ``delattr`` disconnects."""
    return _Port(fget=lambda self: self._ports[name],
                 fset=lambda self, emitter: self._ports[name].assign(emitter),
                 fdel=lambda self: self._ports[name].expel())


class ExplicitChannels(Signal, abc.ABC):
    @state
    class State(Signal.State):
        channels: int = attr.ib(validator=attrs.validators.ge(1), default=1)


class ExplicitChannelsEmitter(ExplicitChannels, Emitter, abc.ABC):
    @state
    class State(ExplicitChannels.State, Emitter.State):
        pass

    @property
    def channels(self) -> int:
        return self._state.channels


class ImplicitChannels(Receiver, Emitter, abc.ABC):

    @property
    def channels(self) -> int:
        """The one non-unit channel count among the inputs (reference :396-406)."""
        counts = {inp.channels for inp in self.inputs_by_port.values()}
        if len(counts) > 1:
            counts.discard(1)
        if len(counts) != 1:
            raise ValueError(f'expected exactly one channel count among inputs of {self.cls_name()!r}, got {sorted(counts)}')
        return next(iter(counts))


class PassThroughResult(ImplicitChannels, abc.ABC):
    input = port('input')

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.PASSTHRU

    def _get_result(self, request: Request) -> np.ndarray:
        return super()._get_result(request) if self._state.enabled else self.input.forward(request)


class NotCached(RuntimeError):
    pass


class BlockCachingEmitter(Emitter, abc.ABC):
    """Kept for API compatibility (reference :424-457).  The reference memoises up to 16 blocks per
    node to de-duplicate fan-out and the overlapping context requests of its filters; the compiled
    plan evaluates every node once per render, so there is nothing to cache."""

    def __init__(self):
        super().__init__()
        self._max_cached_blocks = 16
