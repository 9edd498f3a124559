"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (noah-aviel-dove/signals) from
``/root/reference/src`` behind inert stand-ins for the GUI/audio packages that
are absent from this image, so that its own numpy/scipy render can be executed
to mint golden vectors (``oracle/make_golden.py``) and to validate the numpy
restatement (``oracle/np_oracle.py``).

The reference lives only in the build container; nothing on the GPU box may
call this module (``available()`` is False there).

None of the stubs touches arithmetic.  What is stubbed and why
(reference file:line):
  PyQt5.*            src/signals/__init__.py:7-9, src/signals/ui/theme.py:4-8
  more_itertools.one src/signals/chain/__init__.py:406
  sounddevice        src/signals/chain/dev.py:10, chain/discovery.py:8
  soundfile          src/signals/chain/files.py:8
  matplotlib(.pyplot) src/signals/chain/vis.py:5
  bijection          src/signals/map/__init__.py:10
  numpy.float        src/signals/chain/fx.py:99 (reference pins numpy 1.23.0)
"""
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get('SIGNALS_REFERENCE_SRC', '/root/reference/src')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'signals', 'chain'))


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Dummy


class _Dummy(metaclass=_DummyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Dummy()


def _dummy_module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith('__'):
            raise AttributeError(attr)
        cls = _DummyMeta(attr, (_Dummy,), {})
        setattr(mod, attr, cls)
        return cls

    mod.__getattr__ = _getattr
    return mod


class FakeOutputStream:
    """Headless sounddevice.OutputStream: ``run(n_blocks, frames)`` drives the callback."""
    samplerate = 48000.0
    blocksize = 512

    def __init__(self, device=None, callback=None, channels=1, samplerate=None, **kw):
        self.callback = callback
        self.channels = channels
        if samplerate:
            self.samplerate = float(samplerate)
        self.active = False
        self.captured = []

    def start(self):
        self.active = True

    def stop(self):
        self.active = False

    def close(self):
        self.active = False

    def run(self, n_blocks: int, frames: int):
        for _ in range(n_blocks):
            out = np.zeros((frames, self.channels), dtype=np.float32)
            self.callback(out, frames, None, 0)
            self.captured.append(out)
        return np.concatenate(self.captured)

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, *exc):
        self.close()


FAKE_DEVICE = dict(name='default', index=0, hostapi=0,
                   max_input_channels=2, max_output_channels=2,
                   default_low_input_latency=0.01, default_low_output_latency=0.01,
                   default_high_input_latency=0.1, default_high_output_latency=0.1,
                   default_samplerate=48000.0)


def _install_stubs():
    if 'PyQt5' not in sys.modules:
        pkg = _dummy_module('PyQt5')
        pkg.__path__ = []
        sys.modules['PyQt5'] = pkg
        for sub in ('QtCore', 'QtGui', 'QtWidgets'):
            m = _dummy_module(f'PyQt5.{sub}')
            sys.modules[f'PyQt5.{sub}'] = m
            setattr(pkg, sub, m)
        sys.modules['PyQt5.QtCore'].pyqtSignal = lambda *a, **k: _Dummy()
    if 'more_itertools' not in sys.modules:
        mi = types.ModuleType('more_itertools')

        def one(iterable, too_short=None, too_long=None):
            it = iter(iterable)
            try:
                first = next(it)
            except StopIteration:
                raise too_short or ValueError('too few items in iterable (expected 1)')
            try:
                next(it)
            except StopIteration:
                return first
            raise too_long or ValueError('Expected exactly one item in iterable')

        mi.one = one
        sys.modules['more_itertools'] = mi
    if 'sounddevice' not in sys.modules:
        sd = types.ModuleType('sounddevice')
        sd.OutputStream = FakeOutputStream
        sd.InputStream = FakeOutputStream
        sd.CallbackFlags = int

        class CallbackStop(Exception):
            pass

        sd.CallbackStop = CallbackStop
        sd.query_devices = lambda *a, **k: [dict(FAKE_DEVICE)] if not a else dict(FAKE_DEVICE)
        sys.modules['sounddevice'] = sd
    if 'soundfile' not in sys.modules:
        sys.modules['soundfile'] = _dummy_module('soundfile')
    if 'matplotlib' not in sys.modules:
        mpl = _dummy_module('matplotlib')
        mpl.__path__ = []
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = _dummy_module('matplotlib.pyplot')
        mpl.pyplot = sys.modules['matplotlib.pyplot']
    if 'bijection' not in sys.modules:
        bj = types.ModuleType('bijection')

        class _Inv:
            def __init__(self, owner):
                self._owner = owner

            def __getitem__(self, value):
                return self._owner._inv[id(value)]

            def pop(self, value, *default):
                key = self._owner._inv.pop(id(value), *default)
                if key in self._owner._fwd:
                    del self._owner._fwd[key]
                return key

            def __contains__(self, value):
                return id(value) in self._owner._inv

        class Bijection(dict):
            def __class_getitem__(cls, item):
                return cls

            def __init__(self):
                super().__init__()
                self._fwd = self
                self._inv = {}
                self.inv = _Inv(self)

            def __setitem__(self, k, v):
                if k in self:
                    self._inv.pop(id(dict.__getitem__(self, k)), None)
                dict.__setitem__(self, k, v)
                self._inv[id(v)] = k

            def setdefault(self, k, v):
                if k not in self:
                    self[k] = v
                return dict.__getitem__(self, k)

            def pop(self, k, *default):
                if k in self:
                    v = dict.pop(self, k)
                    self._inv.pop(id(v), None)
                    return v
                if default:
                    return default[0]
                raise KeyError(k)

        bj.Bijection = Bijection
        sys.modules['bijection'] = bj
    if not hasattr(np, 'float'):
        np.float = float  # alias removed in numpy>=1.24; the reference pins 1.23.0


_loaded = None


def load():
    """Return the reference's modules as a namespace (chain, osc, fx, fixed, shape)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference sources not found under {REFERENCE_SRC}')
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import signals.chain as chain
    import signals.chain.fixed as fixed
    import signals.chain.fx as fx
    import signals.chain.osc as osc
    import signals.chain.shape as shape
    ns = types.SimpleNamespace(chain=chain, osc=osc, fx=fx, fixed=fixed, shape=shape)
    _loaded = ns
    return ns


def make_root(ref):
    """A minimal Receiver with one ``input`` port: the pull root (chain/dev.py:173 stand-in)."""
    chain = ref.chain

    class Root(chain.Receiver):
        input = chain.port('input')

        @classmethod
        def flags(cls):
            return super().flags()

    return Root()


def fixed(ref, value):
    f = ref.fixed.Fixed()
    f.get_state().value = np.array(value, ndmin=2, dtype=float)
    return f


def render(ref, emitter, position: int, frames: int, channels: int, rate: int = 48000):
    """Pull one block through the reference's own recursion (chain/__init__.py:296-300)."""
    root = make_root(ref)
    root.input = emitter
    loc = ref.chain.BlockLoc(position=position, rate=rate,
                             shape=ref.chain.Shape(frames=frames, channels=channels))
    out = root.input.request(loc)
    del root.input
    return np.array(out, dtype=np.float64)
