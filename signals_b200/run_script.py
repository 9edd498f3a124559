"""Run one of the reference's entry points (``scripts/*.py``) unchanged on the B200 path:

    python -m signals_b200.run_script /path/to/scripts/edited_sine.py [--blocks N] [--stdin TEXT] [script args...]

* ``sounddevice`` is the headless shim unless the real package imports (no PortAudio in this image);
* the reference's package name ``signals`` (and ``signals.chain.*``, ``signals.map.control``) is aliased to
  this package's mirrors, so ``signals.chain.dev.SinkDevice(...).start()`` pulls blocks through
  ``sigb_render_host``;
* the script's source is executed as is (``runpy``); ``input()`` reads ``--stdin``.
"""
import argparse
import importlib
import io
import runpy
import sys


def alias_signals() -> None:
    import signals_b200
    sys.modules.setdefault('signals', signals_b200)
    for name in ('chain', 'chain.dev', 'chain.discovery', 'chain.fixed', 'chain.osc', 'chain.fx', 'chain.shape',
                 'chain.vis', 'chain.files', 'chain.ext', 'map', 'map.control'):
        mod = importlib.import_module('signals_b200.' + name)
        sys.modules.setdefault('signals.' + name, mod)


def run(path: str, script_args=(), blocks: int = 8, blocksize: int = 512, stdin_text: str = 'default\n'):
    from signals_b200 import sounddevice_shim
    sd = sounddevice_shim.install(blocks=blocks, blocksize=blocksize)
    alias_signals()
    old_argv, old_stdin = sys.argv, sys.stdin
    sys.argv = [path, *script_args]
    sys.stdin = io.StringIO(stdin_text)
    try:
        runpy.run_path(path, run_name='__main__')
    except SystemExit as e:
        if e.code not in (None, 0):
            raise
    finally:
        sys.argv, sys.stdin = old_argv, old_stdin
    return getattr(sd, 'streams', [])


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('script')
    ap.add_argument('--blocks', type=int, default=8)
    ap.add_argument('--blocksize', type=int, default=512)
    ap.add_argument('--stdin', default='default\n')
    args, rest = ap.parse_known_args()
    streams = run(args.script, rest, args.blocks, args.blocksize, args.stdin)
    for i, s in enumerate(streams):
        audio = s.audio()
        print(f'stream {i}: {audio.shape[0]} frames x {audio.shape[1]} ch @ {s.samplerate:g} Hz, peak {abs(audio).max() if audio.size else 0:.4f}')


if __name__ == '__main__':
    main()
