"""File nodes (mirror of signals.chain.files, /root/reference/src/signals/chain/files.py:70-102).

* ``FileReader`` (files.py:70-87) reads ``frames`` frames at the request's position.  Here the file is decoded
  once on the host and lives in HBM as the sample table of a ``Buffer`` node, addressed by absolute frame
  position; frames past the end read as zeros.
* ``FileWriter`` (files.py:89-102) is a pass-through recorder: its audio result is its input.  The plan compiler
  lowers it to its input; the block it would have written is delivered by ``Engine.serve_taps`` (a separate
  render of the tap's input, device -> host, then a positioned write), which ``SinkDevice`` and
  ``Patch.render(taps=True)`` call.

The codec is ``signals_b200.wavio`` (RIFF/WAVE; ``soundfile`` is not in this image).
"""
import attr
import numpy as np

from signals_b200 import SignalFlags, wavio
from signals_b200.chain import PassThroughResult, port, state
from signals_b200.chain.ext import Buffer


class FileReader(Buffer):

    @state
    class State(Buffer.State):
        path: str = attr.ib(default='/dev/null')

    def __init__(self):
        super().__init__()
        self._loaded_path = None
        self.file_rate = None

    def _load(self) -> None:
        path = self.get_state().path
        if path != self._loaded_path:
            if wavio.exists(path):
                samples, self.file_rate = wavio.read(path)
            else:
                samples, self.file_rate = np.zeros((0, 1), dtype=np.float32), None
            self._samples = samples
            self._loaded_path = path

    @property
    def samples(self):
        self._load()
        return self._samples

    @samples.setter
    def samples(self, value) -> None:          # Buffer.__init__ / set_samples
        self._samples = value

    @property
    def channels(self) -> int:
        return int(self.samples.shape[1])


class FileWriter(PassThroughResult):
    input = port('input')

    @state
    class State(PassThroughResult.State):
        path: str = attr.ib(default='/dev/null')

    def __init__(self):
        super().__init__()
        self._writer = None

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.RECORDER

    def deliver(self, position: int, rate: int, block: np.ndarray) -> None:
        """What ``_write`` does in the reference (files.py:96-98): (re)open for this rate / channel count, seek
        to the request's position, write the block."""
        path = self.get_state().path
        if not path or path == '/dev/null':
            return
        w = self._writer
        if w is None or w.path != path or w.rate != rate or w.channels != block.shape[1]:
            self._close()
            w = self._writer = wavio.WavWriter(path, rate, block.shape[1])
        w.write(position, block)

    def _close(self) -> None:
        if self._writer is not None:
            self._writer.close()
            self._writer = None

    def destroy(self) -> None:
        self._close()
        super().destroy()
