"""Minimal RIFF/WAVE codec for the file nodes (`soundfile`, which the reference uses in chain/files.py:8, is
not in this image): float32 writer that can be appended to and sought in, and a reader for PCM 16/24/32-bit
and IEEE float 32/64-bit files.  Pure host-side I/O; no DSP."""
from __future__ import annotations

import os
import struct

import numpy as np

_FMT_PCM, _FMT_FLOAT, _FMT_EXTENSIBLE = 1, 3, 0xFFFE


class WavWriter:
    """Streaming float32 writer: frames are written at their absolute position (gaps read back as zeros)."""

    HEADER = 44

    def __init__(self, path: str, rate: int, channels: int):
        self.path, self.rate, self.channels = path, int(rate), int(channels)
        self.frames = 0
        self._f = open(path, 'wb+')
        self._write_header()

    def _write_header(self) -> None:
        data = self.frames * self.channels * 4
        self._f.seek(0)
        self._f.write(b'RIFF' + struct.pack('<I', 36 + data) + b'WAVE')
        self._f.write(b'fmt ' + struct.pack('<IHHIIHH', 16, _FMT_FLOAT, self.channels, self.rate,
                                            self.rate * self.channels * 4, self.channels * 4, 32))
        self._f.write(b'data' + struct.pack('<I', data))

    def write(self, position: int, block: np.ndarray) -> None:
        block = np.ascontiguousarray(np.broadcast_to(block, (block.shape[0], self.channels)), dtype='<f4')
        self._f.seek(self.HEADER + position * self.channels * 4)
        self._f.write(block.tobytes())
        self.frames = max(self.frames, position + block.shape[0])
        self._write_header()
        self._f.flush()

    def close(self) -> None:
        if self._f:
            self._write_header()
            self._f.close()
            self._f = None


def read(path: str) -> tuple[np.ndarray, int]:
    """Returns ``(samples float32 (frames, channels), rate)``."""
    with open(path, 'rb') as f:
        raw = f.read()
    if raw[:4] != b'RIFF' or raw[8:12] != b'WAVE':
        raise ValueError(f'{path}: not a RIFF/WAVE file')
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        tag, size = raw[pos:pos + 4], struct.unpack('<I', raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if tag == b'fmt ':
            fmt = struct.unpack('<HHIIHH', body[:16])
            if fmt[0] == _FMT_EXTENSIBLE and len(body) >= 26:
                fmt = (struct.unpack('<H', body[24:26])[0],) + fmt[1:]
        elif tag == b'data':
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f'{path}: missing fmt or data chunk')
    kind, channels, rate, _, _, bits = fmt
    if kind == _FMT_FLOAT and bits in (32, 64):
        x = np.frombuffer(data, dtype='<f4' if bits == 32 else '<f8').astype(np.float32)
    elif kind == _FMT_PCM and bits == 16:
        x = np.frombuffer(data, dtype='<i2').astype(np.float32) / 32768.0
    elif kind == _FMT_PCM and bits == 32:
        x = np.frombuffer(data, dtype='<i4').astype(np.float32) / 2147483648.0
    elif kind == _FMT_PCM and bits == 24:
        b = np.frombuffer(data[:len(data) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (np.where(v >= 1 << 23, v - (1 << 24), v)).astype(np.float32) / 8388608.0
    else:
        raise ValueError(f'{path}: unsupported WAVE format {kind} / {bits} bits')
    frames = x.size // channels
    return x[:frames * channels].reshape(frames, channels), int(rate)


def exists(path: str) -> bool:
    return bool(path) and os.path.isfile(path)
