"""Audio devices (mirror of signals.chain.dev, /root/reference/src/signals/chain/dev.py).

``SinkDevice`` is the CALLER of the hot path: its PortAudio callback builds a ``BlockLoc`` and pulls
``self.input.request(loc)`` (dev.py:167-179).  Here the pull is one ``sigb_render_host`` straight into the
callback's ``outdata`` buffer; everything else (stream lifetime, seek/tell, error handling) is as in the
reference.  Uses the real ``sounddevice`` when importable, else the headless shim
(``signals_b200.sounddevice_shim``).
"""
from __future__ import annotations

import abc
import sys
import traceback
import typing

import attr
import attrs.validators
import numpy as np

from signals_b200 import SignalFlags
from signals_b200.chain import (BlockLoc, ChainLayerError, ExplicitChannels, Receiver, Shape, Signal, port, state)


def _sd():
    try:
        import sounddevice as sd
        if not callable(getattr(sd, 'query_devices', None)) or not isinstance(getattr(sd, 'CallbackStop', None), type):
            raise ImportError('not a usable sounddevice module')
    except (ImportError, OSError):
        from signals_b200 import sounddevice_shim as sd
    return sd


class BadPlaybackState(ChainLayerError):
    pass


@attr.s(auto_attribs=True, frozen=True, kw_only=True, order=False)
class DeviceInfo:
    name: str
    index: int
    hostapi: int
    max_input_channels: int
    max_output_channels: int
    default_low_input_latency: float
    default_low_output_latency: float
    default_high_input_latency: float
    default_high_output_latency: float
    default_samplerate: float

    @property
    def is_source(self) -> bool:
        return self.max_input_channels > 0

    @property
    def is_sink(self) -> bool:
        return self.max_output_channels > 0

    def __str__(self) -> str:
        return (f'{self.index:<3} {self.name} ({self.hostapi})\n'
                f'\tMaximum supported channels (I/O): {self.max_input_channels}/{self.max_output_channels}\n'
                f'\tDefault samplerate: {self.default_samplerate}')

    def __lt__(self, other: 'DeviceInfo') -> bool:
        return self.index < other.index


class Device(Signal, abc.ABC):

    def __init__(self, info: DeviceInfo):
        super().__init__()
        self.info = info

    def log(self, msg: typing.Any) -> None:
        print(msg, file=sys.stderr)


class SinkDevice(Device, Receiver, ExplicitChannels):
    input = port('input')

    def __init__(self, info: DeviceInfo):

        @state
        class State(ExplicitChannels.State):
            # the reference validates against max_input_channels (dev.py:101-102); a sink's limit is its outputs
            channels: int = attr.ib(default=1, validator=attrs.validators.in_(range(1, max(info.max_output_channels, 1) + 1)))

        self.State = State
        super().__init__(info=info)
        self.frame_position = 0
        self._stream = None

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.SINK_DEVICE

    def destroy(self) -> None:
        if self.is_open:
            self.close()
        super().destroy()

    @property
    def is_open(self) -> bool:
        return self._stream is not None

    @property
    def is_active(self) -> bool:
        return self.is_open and bool(self._stream.active)

    def open(self) -> None:
        if self.is_open:
            raise BadPlaybackState('The output stream is already open')
        self._stream = _sd().OutputStream(device=self.info.index, callback=self._callback, channels=self._state.channels)

    def close(self) -> None:
        if not self.is_open:
            raise BadPlaybackState('The output stream is not open')
        self._stream.close()
        self._stream = None

    def start(self) -> None:
        if not self.is_open:
            self.open()
        self._stream.start()

    def stop(self) -> None:
        if not self.is_active:
            raise BadPlaybackState('The output stream is not active')
        self._stream.stop()

    def seek(self, position: int) -> None:
        self.frame_position = position * self._stream.blocksize

    def tell(self) -> int:
        return self.frame_position // self._stream.blocksize

    def render_block(self, outdata: np.ndarray, frames: int, rate: int) -> None:
        """The body of the callback: one block request, delivered into ``outdata[:, :channels]``."""
        channels = self._state.channels
        loc = BlockLoc(position=self.frame_position, shape=Shape(channels=channels, frames=frames), rate=rate)
        bound = self._ports['input']
        if bound and outdata.dtype == np.float32 and outdata.strides[1] == 4:
            # fast path (dev.py:173 + :178 in one call): the plan is kept while the graph epoch stands still, the block
            # is ONE captured CUDA graph launch into page-locked staging and one copy into the device's buffer
            # (sigb_render_block); taps get their blocks from that same launch
            from signals_b200 import engine
            eng = engine.default_engine()
            compiled = eng.plan_for(bound.sig, channels, rate, frames)
            block = outdata[:frames, :channels]
            compiled.render_block(loc.position, frames, block)
            if compiled.records.taps:
                eng.serve_taps(bound.sig, loc, rendered=block)
        else:
            outdata[:, :channels] = self.input.request(loc)
        self.frame_position += frames

    def _callback(self, outdata: np.ndarray, frames: int, time: typing.Any, status) -> None:
        if status:
            self.log(status)
        try:
            self.render_block(outdata, frames, int(self._stream.samplerate))
        except Exception:
            self.log(traceback.format_exc())
            raise _sd().CallbackStop
