"""CPU: host-side logic added in round 2 -- the graph epoch behind the engine's dirty flag, tap records in the lowered
plan, the C ABI entry points of the realtime path (no compute without a GPU), and the bench workloads."""
import ctypes
import os

import numpy as np
import pytest

from oracle import cases
from signals_b200 import _lib, chain, engine as engine_mod, plan as plan_mod
from signals_b200.chain import vis

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RATE = 48000


def test_graph_epoch_moves_on_every_observable_edit(ns):
    e0 = chain.graph_epoch()
    f = cases.fixed(ns, [[440.0]])                     # assigning State.value
    assert chain.graph_epoch() > e0
    e1 = chain.graph_epoch()
    osc = ns.Sine()
    assert chain.graph_epoch() == e1                   # constructing a node edits no graph
    osc.hertz = f                                      # port connected
    e2 = chain.graph_epoch()
    assert e2 > e1
    osc.get_state().enabled = False                    # state attribute assigned
    e3 = chain.graph_epoch()
    assert e3 > e2
    osc.set_state(type(osc.get_state())())             # state replaced
    e4 = chain.graph_epoch()
    assert e4 > e3
    del osc.hertz                                      # port disconnected
    assert chain.graph_epoch() > e4
    e5 = chain.graph_epoch()
    f.get_state().value[0, 0] = 1.0                    # in-place write: no setter, no epoch (the engine compares snapshots)
    assert chain.graph_epoch() == e5


def test_plan_is_kept_until_the_graph_is_edited(ns):
    eng = engine_mod.Engine()
    hz = cases.fixed(ns, [[500.0, 600.0]])
    osc = ns.Sine()
    osc.hertz = hz
    g = cases.gain(ns, osc, [[0.2, 0.3]])
    a = eng.plan_for(g, 2, RATE, 512)
    assert eng.plan_for(g, 2, RATE, 512) is a and a.records.tracked
    assert [np.array_equal(arr, snap) for _, arr, snap in a.records.fixed_values] == [True, True]
    hz.get_state().value[0, 1] = 601.0                 # in-place edit of a Fixed: caught by the snapshot compare
    b = eng.plan_for(g, 2, RATE, 512)
    assert b is not a and a.handle is None             # recompiled, the old plan closed
    assert eng.plan_for(g, 2, RATE, 512) is b
    osc.get_state().enabled = False                    # attribute edit -> epoch
    c = eng.plan_for(g, 2, RATE, 512)
    assert c is not b
    other = cases.gain(ns, ns.Square(), [[1.0]])       # an edit of an unrelated graph also moves the epoch:
    d = eng.plan_for(g, 2, RATE, 512)                  # conservative, recompiles once, never stale
    assert d is not c and eng.plan_for(g, 2, RATE, 512) is d
    del other
    eng.clear()


def test_interior_taps_become_tap_records_and_root_taps_pass_through(ns):
    src = cases.osc(ns, 'Sawtooth', [[220.0, 330.0]])
    t1 = vis.Wave()
    t1.input = src
    lp = cases.lowpass(ns, t1, [[900.0, 1500.0]])
    t2 = vis.Wave()
    t2.input = lp
    rec = plan_mod.lower(t2, 2, RATE, 256)
    kinds = [n.kind for n in rec.nodes]
    assert kinds.count(_lib.NODE_TAP) == 1             # only the interior tap; the root tap is the rendered block itself
    assert [(type(t).__name__, creq, idx) for t, creq, idx in rec.taps] == [('Wave', 2, None), ('Wave', 2, 0)]
    assert rec.nodes[rec.root].kind == _lib.NODE_FILTER
    compiled = engine_mod.Engine().compile(t2, 2, RATE, 256)
    d = compiled.describe()
    assert [l['kind'] for l in d['launches']] == ['chain', 'chain'] and d['buffers'] == 1     # the tap keeps the oscillator block aside
    assert _lib.lib().sigb_plan_tap_count(compiled.handle) == 1
    ch = ctypes.c_int32()
    assert _lib.lib().sigb_plan_read_tap(compiled.handle, 0, None, 0, ctypes.byref(ch)) == _lib.SIGB_OK and ch.value == 2
    assert _lib.lib().sigb_plan_read_tap(compiled.handle, 1, None, 0, None) == _lib.SIGB_EINVAL
    compiled.close()
    # a tap record cannot be the root of a plan
    bad = (_lib.SigbNode * 2)()
    bad[0].kind, bad[0].channels = _lib.NODE_ZERO, 1
    bad[1].kind, bad[1].channels = _lib.NODE_TAP, 1
    for k in range(3):
        bad[0].inputs[k] = -1
        bad[1].inputs[k] = 0 if k == 0 else -1
    h = ctypes.c_void_p()
    assert _lib.lib().sigb_plan_create(bad, 2, 1, None, 0, 1, RATE, ctypes.byref(h)) == _lib.SIGB_EINVAL


def test_realtime_entry_points_refuse_without_a_device(ns, engine):
    """No CPU fallback: the block path fails loudly (SIGB_ECUDA -> RuntimeError) when there is no GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    compiled = engine.compile(cases.CASES_BY_NAME['lowpass_c2_8v'].build(ns), 8, RATE)
    out = np.zeros((128, 8), np.float32)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        compiled.render_block(0, 128, out)
    assert compiled.graph_launches == 0
    for key in ('rt_graph', 'rt_max_bytes', 'blockwise_reference', 'voices_pieces', 'reg_variant', 'bank_unroll'):
        compiled.set_option(key, 1)
    with pytest.raises(ValueError):
        compiled.set_option('no_such_option', 1)
    compiled.close()


def test_modulated_cutoff_plans_carry_their_design_launches(ns, engine):
    d = engine.compile(cases.CASES_BY_NAME['lfo_cutoff_cascade'].build(ns), 2, RATE).describe()
    (launch,) = d['launches']
    assert launch['sections'] == 4 and launch['modulated_cutoffs'] == 4 and launch['warm_rows'] == -1   # horizon: per request
    assert d['modulated_parameters'] > 0


def test_bench_workloads_build_and_lower(ns):
    import bench
    args = bench.parse_args.__globals__['argparse'].Namespace(voices=64, seconds=0.01, slab_seconds=10.0)
    for name in ('c2', 'c2m', 'c3', 'c5'):
        wl = bench.WORKLOADS[name](2048 if name in ('c3', 'c5') else 64, 0.01, 0, 1)
        rec = plan_mod.lower(wl.build(ns), wl.out_channels, RATE, wl.frames)
        assert rec.channels == wl.out_channels and wl.units_per_step() > 0
        jobs, units, sample = wl.cpu_sample(1)
        assert jobs and units > 0 and sample
    jobs, units, sample = bench.C2(64, 0.01, 0, 1).cpu_sample_blockwise(2)
    assert len(jobs) == 2 and jobs[0][0] == 'blockwise' and units == 16 * RATE
    del args


def test_delta_form_sections_hold_the_float32_budget(tmp_path):
    """The delta-form sections of k_cascade_delta / k_osc_delta / k_voices, restated in plain C with fmaf
    (tools/delta_form_sim.c, tools/delta_form_sim_hp.c): 8-section cascades in float32 against the float64 state-variable
    cascade on noise + a 7 Hz tone, cutoffs from 5 Hz to 23 kHz -- every case far inside the 1e-4 bar for cascaded IIR, and the
    states equal the state-variable section's own up to float32 rounding (the pointwise hand-over between kernels)."""
    import re
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    for src in ('delta_form_sim.c', 'delta_form_sim_hp.c'):
        exe = tmp_path / src.replace('.c', '')
        subprocess.check_call([gcc, '-O2', '-ffp-contract=off', '-o', str(exe), os.path.join(ROOT, 'tools', src), '-lm'])
        out = subprocess.check_output([str(exe), '5'], text=True)
        rows = re.findall(r'cut0 (\S+): max\|y\| (\S+)\s+err svf32 (\S+)\s+err delta32 (\S+)\s+state-identity dev (\S+)', out)
        assert len(rows) == 8, out
        for cut0, peak, e_svf, e_delta, dev in rows:
            assert float(e_delta) <= 2e-5, (src, cut0, e_delta)
            assert float(e_delta) <= 8.0 * float(e_svf) + 1e-6, (src, cut0, e_svf, e_delta)     # same class as the state-variable form
            if float(cut0) < 20000:
                assert float(dev) <= 1e-5, (src, cut0, dev)
