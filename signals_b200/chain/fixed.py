"""Constant / parameter source (mirrors signals.chain.fixed, /root/reference/src/signals/chain/fixed.py).

In the compiled plan a ``Fixed`` is not a launch: its ``(1, C)`` row becomes a per-channel
parameter table (hertz, phase, cutoff, gain, mix ...) of the kernels that consume it."""
import attr
import numpy as np

from signals_b200 import SignalFlags
from signals_b200.chain import BadStateValue, Emitter, Request, Shape, state, _on_state_setattr


def _validate_array(instance, attribute, new_value):
    if not (isinstance(new_value, np.ndarray) and new_value.ndim == 2):
        raise BadStateValue(instance, attribute.name, new_value, 'must be a 2D array')


class Fixed(Emitter):
    @state
    class State(Emitter.State):
        value: np.ndarray = attr.ib(factory=Emitter.empty_result,
                                    validator=_validate_array,
                                    on_setattr=[attr.setters.validate, _on_state_setattr])

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags()

    @property
    def channels(self) -> int:
        return Shape.of_array(self._state.value).channels

    def _eval(self, request: Request) -> np.ndarray:
        # the state array itself, whatever the request (fixed.py:38-39); no arithmetic, no GPU
        return self._state.value
