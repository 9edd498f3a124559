"""File nodes (mirror of signals.chain.files, /root/reference/src/signals/chain/files.py:70-102).

``FileWriter`` is a pass-through recorder on the audio path (files.py:89-102): the plan compiler lowers it
to its input.  Writing the file needs ``soundfile`` (absent here) and is SURVEY 8f "next"; ``FileReader`` is
served by ``signals_b200.chain.ext.Buffer`` (an HBM-resident sample source).
"""
import attr

from signals_b200 import SignalFlags
from signals_b200.chain import PassThroughResult, port, state


class FileWriter(PassThroughResult):
    input = port('input')

    @state
    class State(PassThroughResult.State):
        path: str = attr.ib(default='')

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.RECORDER
