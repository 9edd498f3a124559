"""Headless stand-in for the `sounddevice` package (PortAudio is absent from this image and from the GPU
box).  It reproduces the slice of the API the reference uses -- ``OutputStream(device, channels, callback,
samplerate, blocksize)`` with ``start/stop/close/active`` and the context-manager protocol,
``query_devices``, ``CallbackStop``, ``CallbackFlags``, ``sleep`` (/root/reference/src/signals/chain/dev.py:
139-179, chain/discovery.py:102, scripts/example_sine.py:44-60) -- and drives the callback synchronously:
``start()`` calls ``callback(outdata, frames, time, status)`` ``BLOCKS`` times with a float32 ``outdata`` of
``blocksize`` frames and keeps every block in ``stream.recorded``.  No arithmetic happens here.

Install it with ``signals_b200.sounddevice_shim.install()`` before importing code that does
``import sounddevice as sd`` (``python -m signals_b200.run_script`` does).
"""
from __future__ import annotations

import sys
import types
import typing

import numpy as np

BLOCKS = 8            # callbacks per start(); override with install(blocks=...)
BLOCKSIZE = 512       # frames per callback when the stream does not fix one (a typical PortAudio choice)
SAMPLERATE = 48000.0

_DEVICE = dict(name='default', index=0, hostapi=0, max_input_channels=2, max_output_channels=64,
               default_low_input_latency=0.01, default_low_output_latency=0.01,
               default_high_input_latency=0.04, default_high_output_latency=0.04,
               default_samplerate=SAMPLERATE)

streams: list = []     # every OutputStream created, in order (scripts keep theirs in a local)


class CallbackStop(Exception):
    pass


class CallbackAbort(Exception):
    pass


class PortAudioError(Exception):
    pass


class CallbackFlags:
    def __init__(self, flags: int = 0):
        self.flags = flags

    def __bool__(self) -> bool:
        return bool(self.flags)


class DeviceList(list):
    def __repr__(self) -> str:
        return '\n'.join(f'{"*" if i == 0 else " "} {d["index"]} {d["name"]}' for i, d in enumerate(self))


def query_devices(device=None, kind=None):
    if device is None and kind is None:
        return DeviceList([dict(_DEVICE)])
    return dict(_DEVICE)


def sleep(msec: float) -> None:
    pass


class OutputStream:
    def __init__(self, samplerate=None, blocksize=None, device=None, channels=None, dtype='float32', latency=None,
                 callback: typing.Optional[typing.Callable] = None, finished_callback=None, **_):
        self.samplerate = float(samplerate or SAMPLERATE)
        self.blocksize = int(blocksize or BLOCKSIZE)
        self.device = device
        self.channels = int(channels or 1)
        self.dtype = dtype
        self.callback = callback
        self.finished_callback = finished_callback
        self.active = False
        self.closed = False
        self.recorded: list[np.ndarray] = []
        streams.append(self)

    def start(self) -> None:
        self.active = True
        try:
            for _ in range(BLOCKS):
                outdata = np.zeros((self.blocksize, self.channels), dtype=self.dtype)
                try:
                    self.callback(outdata, self.blocksize, None, CallbackFlags())
                except CallbackStop:
                    self.recorded.append(outdata)
                    break
                self.recorded.append(outdata)
        finally:
            self.active = False
            if self.finished_callback:
                self.finished_callback()

    def stop(self) -> None:
        self.active = False

    def abort(self) -> None:
        self.active = False

    def close(self) -> None:
        self.active = False
        self.closed = True

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def audio(self) -> np.ndarray:
        """Everything the callbacks produced, concatenated: (frames, channels) float32."""
        return np.concatenate(self.recorded) if self.recorded else np.zeros((0, self.channels), dtype=self.dtype)


class InputStream(OutputStream):
    pass


def install(blocks: typing.Optional[int] = None, blocksize: typing.Optional[int] = None, force: bool = False) -> types.ModuleType:
    """Register this module as ``sounddevice`` (unless the real package is importable and ``force`` is off)."""
    global BLOCKS, BLOCKSIZE
    if blocks is not None:
        BLOCKS = int(blocks)
    if blocksize is not None:
        BLOCKSIZE = int(blocksize)
    me = sys.modules[__name__]
    if not force:
        try:
            import sounddevice  # noqa: F401
            return sys.modules['sounddevice']
        except (ImportError, OSError):
            pass
    sys.modules['sounddevice'] = me
    return me
