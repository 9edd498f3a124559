"""Oscillators (mirror signals.chain.osc, /root/reference/src/signals/chain/osc.py:18-62).

Declarative: the waveform is evaluated by the CUDA chain kernels
(signals_b200/csrc/sigb_kernels.cu, ``osc_wave``), phase taken from the absolute frame index in
float64 exactly as ``frame_range / rate * hertz + phase`` (osc.py:32).  ``OscTable`` (osc.py:65-103)
is dead code in the reference ("significantly slower") and is not provided."""
import abc

from signals_b200 import SignalFlags, _lib
from signals_b200.chain import BlockCachingEmitter, ImplicitChannels, port


class Osc(BlockCachingEmitter, ImplicitChannels, abc.ABC):
    hertz = port('hertz')
    phase = port('phase')

    #: SIGB_WAVE_* code of the waveform the kernels evaluate
    wave: int

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.GENERATOR


class Sine(Osc):
    """sin(2*pi*t), osc.py:40-43."""
    wave = _lib.WAVE_SINE


class Square(Osc):
    """sign(0.5 - t mod 1), osc.py:46-49 (0 exactly on the falling edge)."""
    wave = _lib.WAVE_SQUARE


class Sawtooth(Osc):
    """2*((t - 0.5) mod 1) - 1, osc.py:52-55."""
    wave = _lib.WAVE_SAWTOOTH


class Triangle(Osc):
    """(4*(u mod 0.5) - 1) * sign(u mod 1 - 0.5), u = t - 0.25, osc.py:58-62 (-0.0 on the exact trough)."""
    wave = _lib.WAVE_TRIANGLE
