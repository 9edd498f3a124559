"""Run one of the reference's entry points (``scripts/*.py``) unchanged on the B200 path:

    python -m signals_b200.run_script /path/to/scripts/edited_sine.py [--blocks N] [--stdin TEXT] [script args...]

* ``sounddevice`` is the headless shim (no PortAudio in this image; ``--real-audio`` uses the installed package);
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


def alias_signals() -> dict:
    """Point the reference's package name at the mirrors; returns the displaced sys.modules entries."""
    import signals_b200
    displaced = {k: v for k, v in sys.modules.items() if k == 'signals' or k.startswith('signals.')}
    for k in displaced:
        del sys.modules[k]
    sys.modules['signals'] = signals_b200
    for name in ('chain', 'chain.dev', 'chain.discovery', 'chain.fixed', 'chain.osc', 'chain.fx', 'chain.shape',
                 'chain.vis', 'chain.files', 'chain.ext', 'map', 'map.control'):
        sys.modules['signals.' + name] = importlib.import_module('signals_b200.' + name)
    return displaced


def restore_signals(displaced: dict) -> None:
    for k in [k for k in sys.modules if k == 'signals' or k.startswith('signals.')]:
        del sys.modules[k]
    sys.modules.update(displaced)


def run(path: str, script_args=(), blocks: int = 8, blocksize: int = 512, stdin_text: str = 'default\n',
        real_audio: bool = False):
    from signals_b200 import sounddevice_shim
    old_sd = sys.modules.get('sounddevice')
    sd = sounddevice_shim.install(blocks=blocks, blocksize=blocksize, force=not real_audio)
    displaced = alias_signals()
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
        restore_signals(displaced)
        if old_sd is not None:
            sys.modules['sounddevice'] = old_sd
    return getattr(sd, 'streams', [])


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('script')
    ap.add_argument('--blocks', type=int, default=8)
    ap.add_argument('--blocksize', type=int, default=512)
    ap.add_argument('--stdin', default='default\n')
    ap.add_argument('--real-audio', action='store_true', help='use the installed sounddevice/PortAudio instead of the headless shim')
    args, rest = ap.parse_known_args()
    streams = run(args.script, rest, args.blocks, args.blocksize, args.stdin, args.real_audio)
    for i, s in enumerate(streams):
        audio = s.audio()
        print(f'stream {i}: {audio.shape[0]} frames x {audio.shape[1]} ch @ {s.samplerate:g} Hz, peak {abs(audio).max() if audio.size else 0:.4f}')


if __name__ == '__main__':
    main()
