"""Channel reshaping (mirrors signals.chain.shape, /root/reference/src/signals/chain/shape.py).

``Merge`` (shape.py:60-74) works in the reference and is lowered to column-range copies.
``Flatten`` / ``FlattenUnit`` / ``Select`` are broken in the reference (they reduce over frames and
return 1-D arrays that the block cache rejects, shape.py:32-57); the names are kept so patches
load, and rendering them raises.  The working N->1 mixdowns are ``signals_b200.chain.ext``."""
import abc

import attr
import attrs.validators

from signals_b200 import SignalFlags
from signals_b200.chain import BlockCachingEmitter, Receiver, port, state


class Shaper(BlockCachingEmitter, Receiver, abc.ABC):

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.EFFECT


class Scalar(Shaper, abc.ABC):
    input: Receiver.BoundPort = port('input')

    @property
    def channels(self) -> int:
        return 1


class Flatten(Scalar):
    pass


class FlattenUnit(Scalar):
    pass


class Select(Scalar):
    @state
    class State(BlockCachingEmitter.State):
        index: int = attr.ib(validator=attrs.validators.ge(0), default=0)


class Merge(Shaper):
    """hstack(left, right), each side requested at its own channel count (shape.py:60-74)."""
    left: Receiver.BoundPort = port('left')
    right: Receiver.BoundPort = port('right')

    @property
    def channels(self) -> int:
        return sum(inp.channels for inp in self.inputs_by_port.values())
