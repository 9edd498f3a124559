"""signals_b200 -- B200-native block-render engine behind the `signals` node-graph API.

The reference (noah-aviel-dove/signals) renders its node graph by Python recursion over numpy
blocks (src/signals/chain/__init__.py:245-315).  This package keeps that API --
``signals_b200.chain`` mirrors ``signals.chain`` name for name -- and moves only the block render
to hand-written sm_100a CUDA kernels in ``libsigb200.so`` (C ABI in ``include/sigb200.h``).

There is no CPU fallback: rendering without the CUDA library or without a GPU raises.
"""
import enum

PortName = str

__version__ = '0.1.0'


class SignalsError(Exception):
    """Root of the exception tree (mirrors signals.SignalsError, src/signals/__init__.py:18-21)."""

    def __str__(self) -> str:
        return ' '.join((type(self).__name__, *map(str, self.args)))


class SignalFlags(enum.Flag):
    """Node capability flags (mirrors signals.SignalFlags, src/signals/__init__.py:27-58)."""
    CYCLIC = enum.auto()
    SINK_DEVICE = enum.auto()
    SOURCE_DEVICE = enum.auto()
    DEVICE = SINK_DEVICE | SOURCE_DEVICE
    GENERATOR = enum.auto()
    EFFECT = enum.auto()
    AUDIO = GENERATOR | EFFECT | SOURCE_DEVICE
    EPOCH = enum.auto()
    RECORDER = enum.auto()
    VIS = enum.auto()
    PASSTHRU = enum.auto()
    SIDE_EFFECT = VIS | RECORDER | PASSTHRU
