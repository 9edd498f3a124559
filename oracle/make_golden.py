"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

For every case in ``oracle/cases.py`` the graph is built from the reference's own node
classes and pulled through the reference's own recursion (``ref_harness.render``).
The numpy/scipy versions used are recorded in each file because they differ from the
reference's pins (requirements.txt:5-6).
"""
import json
import os
import sys

import numpy as np
import scipy

from oracle import cases, ref_harness

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main(names=None):
    ref = ref_harness.load()
    ns = cases.ref_namespace(ref)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    meta = {'numpy': np.__version__, 'scipy': scipy.__version__, 'python': sys.version.split()[0],
            'reference_pins': {'numpy': '1.23.0', 'scipy': '1.10.1'}, 'cases': {}}
    for case in cases.CASES:
        if names and case.name not in names:
            continue
        graph = case.build(ns)
        if case.block:
            parts = [ref_harness.render(ref, graph, case.position + r, min(case.block, case.frames - r), case.channels, case.rate)
                     for r in range(0, case.frames, case.block)]
            out = np.concatenate([np.broadcast_to(b, (min(case.block, case.frames - r), case.channels))
                                  for b, r in zip(parts, range(0, case.frames, case.block))])
        else:
            out = ref_harness.render(ref, graph, case.position, case.frames, case.channels, case.rate)
        out = np.broadcast_to(out, (case.frames, case.channels))[::case.stride]
        np.savez_compressed(os.path.join(GOLDEN_DIR, f'{case.name}.npz'), out=np.ascontiguousarray(out))
        meta['cases'][case.name] = dict(position=case.position, frames=case.frames, channels=case.channels,
                                        rate=case.rate, stride=case.stride, tol=case.tol, note=case.note, block=case.block,
                                        absmax=float(np.nanmax(np.abs(out))))
        print(f'{case.name:24s} {out.shape} absmax={meta["cases"][case.name]["absmax"]:.6g}')
    errors = {}
    for name, build, frames, channels, exc in cases.ERROR_CASES + cases.RUNTIME_ERROR_CASES:
        try:
            ref_harness.render(ref, build(ns), 0, frames, channels)
            got = None
        except Exception as e:  # noqa: BLE001 - recording whatever the reference raises
            got = type(e).__name__
        errors[name] = got
        print(f'{name:32s} reference raises {got} (expected {exc})')
        assert got == exc, (name, got, exc)
    meta['error_cases'] = errors
    meta_path = os.path.join(GOLDEN_DIR, 'META.json')
    if names and os.path.exists(meta_path):      # minting a subset: merge into the existing record
        with open(meta_path) as f:
            old = json.load(f)
        old['cases'].update(meta['cases'])
        old['error_cases'] = meta['error_cases']
        meta = old
    with open(meta_path, 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main(sys.argv[1:] or None)
