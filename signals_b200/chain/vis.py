"""Visualisation taps (mirror of signals.chain.vis, /root/reference/src/signals/chain/vis.py:19-89).

On the audio path ``Wave`` / ``Spec`` are pass-through nodes (``PassThroughResult``, chain/__init__.py:
409-417): they forward their input unchanged and queue the block for the GUI (vis.py:61-64).  The GUI is
out of scope; the plan compiler lowers them to their input, so a patch with taps renders on the GPU
exactly like the patch without them.  ``Vis.q`` is kept so a host-side consumer can still be attached.
"""
import abc
import queue

import attr

from signals_b200 import SignalFlags
from signals_b200.chain import PassThroughResult, port, state


class Vis(PassThroughResult, abc.ABC):
    input = port('input')

    def __init__(self):
        super().__init__()
        self.q = queue.Queue()

    @classmethod
    def flags(cls) -> SignalFlags:
        return super().flags() | SignalFlags.VIS

    def deliver(self, position: int, rate: int, block) -> None:
        """The reference queues every block it forwards for the GUI thread (vis.py:61-64)."""
        self.q.put(block)


class Wave(Vis):

    @state
    class State(Vis.State):
        min_amp: float = attr.ib(default=-1.0)
        max_amp: float = attr.ib(default=1.0)


class Spec(Vis):
    pass
