"""Binary effects and Butterworth filters (mirror signals.chain.fx,
/root/reference/src/signals/chain/fx.py).

Declarative nodes; arithmetic lives in libsigb200:
  Mix / RingMod / Amp     -> k_ewise                       (fx.py:35-46, 55-60)
  Gain                    -> folded into its producer's chain launch (fx.py:49-52)
  LowPass / HighPass      -> state-variable sections inside the chain launch (fx.py:85-151)
``BandPass`` / ``BandStop`` exist for name compatibility but, as in the reference (whose
``*crit_2[0, i]`` at fx.py:99 raises TypeError), cannot be rendered."""
import abc
import enum

from signals_b200 import SignalFlags, _lib
from signals_b200.chain import BlockCachingEmitter, ImplicitChannels, Receiver, port


class Effect(BlockCachingEmitter, ImplicitChannels, abc.ABC):

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.EFFECT


class BinaryEffect(Effect, abc.ABC):
    left: Receiver.BoundPort = port('left')
    right: Receiver.BoundPort = port('right')


class Mix(BinaryEffect):
    """mix*left + (1-mix)*right with ``mix`` sampled at block rate -- a crossfade (fx.py:35-40)."""
    mix: Receiver.BoundPort = port('mix')


class RingMod(BinaryEffect):
    """left * right (fx.py:43-46)."""


class Gain(BinaryEffect):
    """left * right, ``right`` sampled at block rate (fx.py:49-52)."""


class Amp(BinaryEffect):
    """copysign(left ** right, left), ``right`` at block rate (fx.py:55-60)."""


class CritFilter(Effect, abc.ABC):
    input: Receiver.BoundPort = port('input')

    #: Butterworth order; a class-level knob exactly as in the reference (fx.py:66)
    order = 2

    class Type(enum.StrEnum):
        low_pass = 'lp'
        high_pass = 'hp'
        band_pass = 'bp'
        band_stop = 'bs'

        @property
        def is_band(self) -> bool:
            return self.startswith('b')

    @abc.abstractmethod
    def type(self) -> 'CritFilter.Type':
        raise NotImplementedError

    def context_frames(self) -> int:
        """Zero-state warm-up the reference applies to *every* block (fx.py:82-83, 93-105); the
        B200 plan carries true filter state between contiguous blocks and only warms up after a seek."""
        return 100


class SingleCritFilter(CritFilter, abc.ABC):
    cutoff: Receiver.BoundPort = port('cutoff')


class DoubleCritFilter(CritFilter, abc.ABC):
    low: Receiver.BoundPort = port('low')
    high: Receiver.BoundPort = port('high')


class LowPass(SingleCritFilter):
    subtype = _lib.FILT_LOWPASS

    def type(self) -> CritFilter.Type:
        return self.Type.low_pass


class HighPass(SingleCritFilter):
    subtype = _lib.FILT_HIGHPASS

    def type(self) -> CritFilter.Type:
        return self.Type.high_pass


class BandPass(DoubleCritFilter):

    def type(self) -> CritFilter.Type:
        return self.Type.band_pass


class BandStop(DoubleCritFilter):

    def type(self) -> CritFilter.Type:
        return self.Type.band_stop
