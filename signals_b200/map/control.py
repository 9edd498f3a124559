"""``signals.map.control`` stand-in: a ``Controller`` that can ``load`` a ``.sigs`` patch headlessly
(/root/reference/src/signals/map/control.py:572-594, 705-727) by delegating to ``signals_b200.sigs``."""
import shlex

from signals_b200 import sigs


class Controller:

    def __init__(self, interactive: bool = False, stdout=None, paths=()):
        self.patch = sigs.Patch()
        self.stdout = stdout

    def onecmd(self, line: str) -> None:
        tokens = shlex.split(line, comments=True)
        if tokens and tokens[0] == 'load':
            with open(tokens[1]) as f:
                for patch_line in f:
                    self.patch.execute(patch_line)
        else:
            self.patch.execute(line)
