"""TEST INFRASTRUCTURE -- achieved max-abs error of the CUDA path on every golden case, written as a table for
profiles/ (run on the GPU box; the goldens under tests/golden/ were minted from the unmodified reference), plus the
smoke configuration per kernel variant with its worst voice.

    python tests/parity_report.py > gpurun_out/parity.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from conftest import load_golden, max_abs_err      # noqa: E402
from oracle import cases as case_mod, np_oracle    # noqa: E402
from signals_b200 import engine as engine_mod      # noqa: E402
from signals_b200 import workloads as w            # noqa: E402

RATE = 48000


def main():
    eng = engine_mod.Engine()
    ns = w.b200_namespace()
    print('%-28s %8s %6s %10s %10s  %s' % ('case', 'frames', 'ch', 'max-abs', 'tolerance', 'kernel path'))
    for case in case_mod.CASES:
        c = eng.compile(case.build(ns), case.channels, case.rate, case.frames)
        for k, v in case.options.items():
            c.set_option(k, v)
        if case.block:
            out = np.concatenate([c.render_device(case.position + r, min(case.block, case.frames - r)).cpu().numpy()
                                  for r in range(0, case.frames, case.block)])
        else:
            out = c.render_device(case.position, case.frames).cpu().numpy()
        kinds = ','.join(l['kind'] + (':%ds' % l['sections'] if 'sections' in l else '') for l in c.describe()['launches'])
        c.close()
        got, want = out[::case.stride], load_golden(case.name)
        if case.name == 'amp_frac':
            a = load_golden('amp_int')
            keep = np.stack([np.abs(a[:, 0]) ** 0.5, np.abs(a[:, 1]) ** (1 / 3)], axis=1) > 0.25
            got, want = np.where(keep, got, 0.0), np.where(keep, want, 0.0)
        err = max_abs_err(got, want)
        flag = '' if err <= case.tol else '  <-- ABOVE TOLERANCE'
        print('%-28s %8d %6d %10.3e %10.1e  %s%s' % (case.name, case.frames, case.channels, err, case.tol, kinds, flag))
    # the smoke configuration (C2 in miniature: 256 voices x 0.25 s), per kernel variant, with the worst voice
    print()
    v, frames = 256, 12000
    hertz, phase, cutoff, g = w.voice_params(2, v)
    graph = w.gain(ns, w.lowpass(ns, w.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])
    want = np_oracle.render_voice_chain(0, frames, RATE, hertz, phase, cutoff, g)
    for label, opts in (('seq', dict(force_seq=1)), ('scan9', dict(scan_variant=9)), ('scan16 (f64 carry)', dict(scan_variant=16)),
                        ('scan18 (f32 carry, default)', dict(scan_variant=18)), ('osc_reg', dict(osc_reg=1))):
        c = eng.compile(graph, v, RATE)
        for k, val in opts.items():
            c.set_option(k, val)
        got = c.render_device(0, frames).cpu().numpy()
        c.close()
        e = np.abs(got - want).max(axis=0)
        i = int(e.argmax())
        print('smoke 256x12000 %-28s max-abs %.3e at voice %d (hertz %.1f cutoff %.1f gain %.2f), median voice %.3e'
              % (label, e.max(), i, hertz[i], cutoff[i], g[i], np.median(e)))


if __name__ == '__main__':
    main()
