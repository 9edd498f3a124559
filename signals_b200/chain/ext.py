"""Nodes the reference lacks or ships broken, needed by the BASELINE configs.

* ``GroupSum`` -- N -> G channel mixdown.  The reference's ``Flatten`` cannot run
  (/root/reference/src/signals/chain/shape.py:32-35 sums over frames and returns 1-D), so the
  additive-bank config defines the mixdown as ``x.reshape(F, G, C//G).sum(-1)`` (oracle: numpy).
* ``PanSum``   -- stereo mixdown ``L = sum((1-pan) y), R = sum(pan y)`` for the voice-bank config.
* ``Buffer``   -- an HBM-resident sample source addressed by absolute frame position (the
  ``FileReader`` of chain/files.py:70-87 without the disk); zeros past its end.
"""
import attr
import attrs.validators
import numpy as np

from signals_b200 import SignalFlags
from signals_b200.chain import BlockCachingEmitter, Emitter, Receiver, port, state


class GroupSum(BlockCachingEmitter, Receiver):
    input: Receiver.BoundPort = port('input')

    @state
    class State(BlockCachingEmitter.State):
        groups: int = attr.ib(validator=attrs.validators.ge(1), default=1)

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.EFFECT

    @property
    def channels(self) -> int:
        return self._state.groups


class PanSum(BlockCachingEmitter, Receiver):
    input: Receiver.BoundPort = port('input')
    pan: Receiver.BoundPort = port('pan')

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.EFFECT

    @property
    def channels(self) -> int:
        return 2


class Buffer(BlockCachingEmitter):
    """``samples``: (frames, channels) array-like, uploaded once to HBM as float32 (or a CUDA
    torch tensor used in place)."""

    def __init__(self, samples=None):
        super().__init__()
        self.samples = None
        if samples is not None:
            self.set_samples(samples)

    def set_samples(self, samples) -> None:
        if hasattr(samples, 'is_cuda'):
            if samples.dim() != 2:
                raise ValueError('Buffer samples must be 2-D (frames, channels)')
        else:
            samples = np.asarray(samples)
            if samples.ndim != 2:
                raise ValueError('Buffer samples must be 2-D (frames, channels)')
        self.samples = samples

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.GENERATOR

    @property
    def channels(self) -> int:
        return int(self.samples.shape[1])
