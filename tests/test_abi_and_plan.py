"""CPU: the C-ABI library loads and exports every symbol include/sigb200.h declares; the host
planner (graph -> records -> fused launches) and the reference's error behaviour.  No compute."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.signal

from conftest import ROOT
from oracle import cases
from signals_b200 import _lib, engine as engine_mod, plan as plan_mod
from signals_b200 import chain
from signals_b200.chain import ext, fx, osc, shape


def test_header_symbols_are_exported():
    text = open(os.path.join(ROOT, 'include', 'sigb200.h')).read()
    declared = set(re.findall(r'\b(sigb_[a-z0-9_]+)\s*\(', text))
    assert len(declared) >= 17
    lib = _lib.lib()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert lib.sigb_abi_version() == 2


def test_no_fallback_without_gpu(ns, engine):
    """Without a CUDA device a render raises; it never silently computes on the CPU."""
    if _lib.lib().sigb_device_count() > 0:
        pytest.skip('GPU present')
    g = cases.CASES_BY_NAME['sine_basic'].build(ns)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        engine_mod.render(g, 0, 16, 4)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle or scipy (no CPU fallback)."""
    pkg = os.path.join(ROOT, 'signals_b200')
    pat = re.compile(r'^\s*(import|from)\s+(oracle|scipy)\b', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f
    # tools/ measure the product: they build their graphs from signals_b200.workloads, never from oracle/
    for f in os.listdir(os.path.join(ROOT, 'tools')):
        if f.endswith('.py'):
            assert not pat.search(open(os.path.join(ROOT, 'tools', f)).read()), f
    # bench.py may execute oracle/ only in its CPU legs (cpu_baseline and --impl reference)
    bench = open(os.path.join(ROOT, 'bench.py')).read()
    uses = [m.start() for m in re.finditer(r'^\s*(import|from)\s+oracle\b', bench, re.M)]
    cpu_leg, gpu_leg = bench.index('# CPU side: the oracle port'), bench.index('def run_b200(')
    assert uses and all(cpu_leg < u < gpu_leg for u in uses), 'bench.py imports oracle/ outside its CPU legs'


@pytest.mark.parametrize('case', cases.CASES, ids=lambda c: c.name)
def test_every_case_compiles(case, ns, engine):
    compiled = engine.compile(case.build(ns), case.channels, case.rate, case.frames)
    d = compiled.describe()
    assert d['channels'] == case.channels and d['launches']
    compiled.close()


def test_voice_chain_fuses_into_one_launch(ns, engine):
    """Sine<-Fixed -> LowPass<-Fixed -> Gain<-Fixed (config C2) is ONE chain launch with 1 section."""
    d = engine.compile(cases.CASES_BY_NAME['lowpass_c2_8v'].build(ns), 8, 48000).describe()
    (launch,) = d['launches']
    want = dict(kind='chain', channels=8, source='osc', wave='sine', sections=1, sections_padded=1, gain=True)
    assert {k: launch[k] for k in want} == want
    assert 0 < launch['warm_rows'] < 48000        # decay horizon of the slowest voice (time-split pieces)
    assert d['buffers'] == 0 and d['context'] == 100


def test_cascade_fuses_sections(ns, engine):
    d = engine.compile(cases.CASES_BY_NAME['cascade8'].build(ns), 4, 48000).describe()
    assert len(d['launches']) == 1 and d['launches'][0]['sections'] == 8 and d['context'] == 800
    d = engine.compile(cases.CASES_BY_NAME['highpass_order3'].build(ns), 2, 48000).describe()
    assert d['launches'][0]['sections'] == 2 and d['launches'][0]['sections_padded'] == 2


def test_fanout_materialises_shared_node(ns, engine):
    d = engine.compile(cases.CASES_BY_NAME['fanout'].build(ns), 4, 48000).describe()
    kinds = [l['kind'] for l in d['launches']]
    assert kinds.count('chain') == 3 and d['buffers'] >= 3


@pytest.mark.parametrize('name,build,frames,channels,exc', cases.ERROR_CASES, ids=lambda v: v if isinstance(v, str) else '')
def test_error_cases_raise_like_the_reference(name, build, frames, channels, exc, ns, engine):
    with pytest.raises(Exception) as info:
        engine.compile(build(ns), channels, 48000, frames)
    assert any(k.__name__ == exc for k in type(info.value).__mro__), type(info.value).__mro__
    assert isinstance(info.value, chain.ChainLayerError)


def test_bad_shape_when_block_wider_than_request(ns, engine):
    g = cases.osc(ns, 'Sine', [[1.0, 2.0, 3.0, 4.0]])
    with pytest.raises(chain.BadShape):
        engine.compile(g, 2, 48000, 16)


def test_band_filters_are_unrenderable_like_the_reference(ns, engine):
    f = fx.BandPass()
    f.input = cases.osc(ns, 'Sine', [[100.0]])
    f.low = cases.fixed(ns, [[100.0]])
    f.high = cases.fixed(ns, [[200.0]])
    with pytest.raises(TypeError):
        engine.compile(f, 1, 48000, 16)


def test_unsupported_nodes_raise_at_compile_time(ns, engine):
    f = shape.Flatten()
    f.input = cases.osc(ns, 'Sine', [[100.0, 200.0]])
    with pytest.raises(chain.UnsupportedGraph):
        engine.compile(f, 1, 48000, 16)


def test_modulated_parameters_lower_to_a_parameter_program(ns, engine):
    """An emitter other than Fixed on a block-rate port (osc.py:28-30, fx.py:39,52,59) becomes a float64
    parameter program evaluated once per request; a modulated Gain is not folded into its chain."""
    d = engine.compile(cases.CASES_BY_NAME['lfo_hertz'].build(ns), 2, 48000).describe()
    assert d['modulated_parameters'] >= 3                      # Mix, Gain, and the three LFO oscillators
    assert [l['kind'] for l in d['launches']] == ['chain']
    # a modulated Gain at the end of a run (here: oscillator -> Gain(LFO), feeding a LowPass) rides on that run's gain table,
    # re-derived per request on the device (k_gain_rows); the filter behind it reads the materialised block
    d = engine.compile(cases.CASES_BY_NAME['lfo_chain'].build(ns), 2, 48000).describe()
    assert [l['kind'] for l in d['launches']] == ['chain', 'chain']
    assert d['launches'][1]['source'] == 'block' and d['launches'][1]['gain'] is True
    # a voice bank with a modulated oscillator is not fused with its mix-down
    ps = ext.PanSum()
    ps.input = cases.CASES_BY_NAME['lfo_gain'].build(ns)
    ps.pan = cases.fixed(ns, [[0.2, 0.8]])
    kinds = [l['kind'] for l in engine.compile(ps, 2, 48000).describe()['launches']]
    assert 'voices' not in kinds and kinds[-1] == 'reduce'


def test_modulated_filter_cutoff_plans_on_the_device(ns, engine):
    """A cutoff driven by an emitter is lowered to a parameter-program row plus a per-request device-side design of
    the filter's sections (k_design); such a chain never takes the time-parallel kernels (their tables come from a
    host-side design, warm_rows is unknown), and a cutoff narrower than the request still raises the reference's
    IndexError (fx.py:99)."""
    f = fx.LowPass()
    f.input = cases.osc(ns, 'Sine', [[440.0, 550.0]])
    f.cutoff = cases.osc(ns, 'Sine', [[2.0, 3.0]])
    d = engine.compile(f, 2, 48000, 16).describe()
    (l,) = d['launches']
    assert l['kind'] == 'chain' and l['modulated_cutoffs'] == 1 and l['warm_rows'] == -1 and d['modulated_parameters'] >= 1
    narrow = fx.LowPass()
    narrow.input = cases.osc(ns, 'Sine', [[440.0, 550.0]])
    narrow.cutoff = cases.osc(ns, 'Sine', [[2.0]])
    with pytest.raises(IndexError):
        engine.compile(narrow, 2, 48000, 16)
    # four chained filters with modulated cutoffs are still ONE launch: four sections, four per-request designs
    d = engine.compile(cases.CASES_BY_NAME['lfo_cutoff_cascade'].build(ns), 2, 48000).describe()
    (l,) = d['launches']
    assert l['sections'] == 4 and l['modulated_cutoffs'] == 4 and l['source'] == 'osc'
    # a filter inside a parameter graph has no one-frame meaning here
    car = osc.Sine()
    car.hertz = cases.lowpass(ns, cases.osc(ns, 'Sine', [[3.0]]), [[10.0]])
    with pytest.raises(chain.UnsupportedGraph):
        engine.compile(car, 1, 48000, 16)


def test_api_surface_mirrors_reference():
    s = chain.Shape(frames=10, channels=2)
    assert s == (10, 2) and (1, 1) <= chain.Shape(10, 1) <= s and not (chain.Shape(3, 2) <= s) and (10, 1) <= s
    loc = chain.BlockLoc(position=150, rate=48000, shape=s)
    assert loc.end_position == 160 and loc.resize(1).shape == (1, 2) and loc.reslice(5).shape == (10, 5)
    assert loc.before(100).position == 50 and loc.before(100).shape == (100, 2)
    assert loc.before(1000).position == 0 and loc.before(1000).shape == (150, 2)
    assert loc.after(7).position == 160 and loc.after(7).shape == (7, 2)
    assert loc.resize(4) <= loc and not (loc.after(1) <= loc)
    assert np.array_equal(loc.frame_range, np.arange(150, 160).reshape(-1, 1))
    sine = osc.Sine()
    assert sorted(sine.port_names()) == ['hertz', 'phase'] and sine.inputs_by_port == {}
    f = cases.b200_namespace().Fixed()
    f.get_state().value = np.array([[1.0, 2.0]])
    sine.hertz = f
    assert sine.inputs_by_port == {'hertz': f} and sine.channels == 2 and ('hertz', sine) in f.outputs_with_ports
    del sine.hertz
    assert not sine.hertz and f.outputs_with_ports == set()
    with pytest.raises(chain.BadStateValue):
        f.get_state().value = np.zeros(3)
    with pytest.raises(chain.BadStateSchema):
        f.set_state(sine.get_state())
    assert osc.Sine.cls_name() == 'signals_b200.chain.osc.Sine'
    assert fx.LowPass().type() == 'lp' and fx.LowPass.order == 2 and fx.LowPass().context_frames() == 100
    assert sine.state_attrs() == {'enabled'} and f.state_attrs() == {'enabled', 'value'}


def test_upstream_is_post_order_of_receivers(ns):
    g = cases.CASES_BY_NAME['lowpass_c2_8v'].build(ns)
    names = [type(n).__name__ for n in g.upstream()]
    assert names == ['Sine', 'LowPass', 'Gain']


@pytest.mark.parametrize('subtype,btype', [(_lib.FILT_LOWPASS, 'lp'), (_lib.FILT_HIGHPASS, 'hp')])
@pytest.mark.parametrize('order', [1, 2, 3, 4, 8, 16])
@pytest.mark.parametrize('wn', [100 / 24000, 0.05, 0.5, 0.95])
def test_library_filter_design_equals_scipy_butter(subtype, btype, order, wn):
    """The state-variable sections libsigb200 designs have the transfer function of
    scipy.signal.butter (what chain/fx.py:115-121 calls): compare frequency responses."""
    lib = _lib.lib()
    coef = (ctypes.c_double * (4 * 16))()
    n = lib.sigb_design_butter(subtype, order, wn, coef, 16)
    assert n == order // 2 + order % 2
    w = np.linspace(1e-3, np.pi - 1e-3, 257)
    z = np.exp(1j * w)
    s = (z - 1) / (z + 1)            # bilinear variable (2*fs factor folded into g)
    h = np.ones_like(z)
    for k in range(n):
        g, r2, kind = coef[4 * k], coef[4 * k + 1], int(coef[4 * k + 2])
        sn = s / g
        if kind & 2:
            h *= (sn if kind & 1 else 1) / (sn + 1)
        else:
            h *= (sn * sn if kind & 1 else 1) / (sn * sn + r2 * sn + 1)
    _, want = scipy.signal.sosfreqz(scipy.signal.butter(order, wn, btype, output='sos'), worN=w)
    assert np.abs(h - want).max() < 1e-9


def test_signature_changes_with_state_and_topology(ns):
    g = cases.CASES_BY_NAME['lowpass_c2_8v'].build(ns)
    s0 = plan_mod.signature(g)
    assert plan_mod.signature(g) == s0
    g.inputs_by_port['right'].get_state().value = np.full((1, 8), 0.5)
    s1 = plan_mod.signature(g)
    assert s1 != s0
    g.inputs_by_port['left'].get_state().enabled = False
    assert plan_mod.signature(g) != s1


def test_plan_walker_accepts_reference_objects():
    """Drop-in boundary: the same lowering runs on the reference's OWN node objects."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip('reference sources only exist in the build container')
    ref = ref_harness.load()
    rns = cases.ref_namespace(ref)
    for name in ('lowpass_c2_8v', 'lowpass_test_sigs', 'mix', 'cascade8', 'disabled_osc', 'unconnected'):
        case = cases.CASES_BY_NAME[name]
        a = plan_mod.lower(case.build(rns), case.channels, case.rate, case.frames)
        b = plan_mod.lower(case.build(cases.b200_namespace()), case.channels, case.rate, case.frames)
        assert len(a.nodes) == len(b.nodes) and np.array_equal(a.data, b.data)
        for x, y in zip(a.nodes, b.nodes):
            assert bytes(x) == bytes(y)


# ------------------------------------------------------------------------------------------------
# fused render + mix-down lowering (configs C3 / C5)
# ------------------------------------------------------------------------------------------------

def _set_default(key, value):
    assert _lib.lib().sigb_set_default_option(key.encode(), int(value)) == 0


def test_bank_lowers_to_one_fused_launch(ns, engine):
    hertz, phase, amp = cases.bank_params(3, 2048, 256)
    d = engine.compile(cases.build_bank(ns, ext, hertz, phase, amp, 8), 8, 48000).describe()
    assert d['launches'] == [dict(kind='bank', node=d['launches'][0]['node'], partials=2048, groups=8, gain=True)]
    assert d['buffers'] == 0 and d['state_doubles'] == 0


def test_non_sine_bank_and_filtered_bank_fall_back_to_materialised_blocks(ns, engine):
    gs = ext.GroupSum()
    gs.get_state().groups = 2
    gs.input = cases.osc(ns, 'Square', [[100.0, 200.0, 300.0, 400.0]])
    kinds = [l['kind'] for l in engine.compile(gs, 2, 48000).describe()['launches']]
    assert kinds == ['chain', 'reduce']
    gs.input = cases.lowpass(ns, cases.osc(ns, 'Sine', [[100.0, 200.0, 300.0, 400.0]]), [[500.0] * 4])
    kinds = [l['kind'] for l in engine.compile(gs, 2, 48000).describe()['launches']]
    assert kinds == ['chain', 'reduce']


def test_instances_lower_to_one_voices_launch(ns, engine):
    prm = cases.instance_params(5, 600)
    d = engine.compile(cases.build_instances(ns, ext, prm), 2, 48000).describe()
    (l,) = d['launches']
    assert l['kind'] == 'voices' and l['segments'] == 12 and l['channels'] == 600 and l['channels_per_thread'] == 1
    assert d['context'] == 100          # LowPass / HighPass context (fx.py:82-83) survives the fusion
    # filter state: one section (2 doubles) per filtered instance
    assert d['state_doubles'] == 2 * int((prm['filt'] > 0).sum())


def test_fan_out_below_a_pansum_disables_the_fusion(ns, engine):
    """A chain consumed twice must be materialised once (the block cache's job, chain/__init__.py:424-457)."""
    o = cases.gain(ns, cases.osc(ns, 'Sine', [[100.0, 200.0]]), [[0.5, 0.5]])
    m = ns.Merge()
    m.left = o
    m.right = o
    ps = ext.PanSum()
    ps.input = m
    ps.pan = cases.fixed(ns, [[0.1, 0.2, 0.3, 0.4]])
    kinds = [l['kind'] for l in engine.compile(ps, 2, 48000).describe()['launches']]
    assert 'voices' not in kinds and kinds[-1] == 'reduce'


def test_fuse_reduce_default_option(ns, engine):
    hertz, phase, amp = cases.bank_params(3, 64, 32)
    _set_default('fuse_reduce', 0)
    try:
        kinds = [l['kind'] for l in engine.compile(cases.build_bank(ns, ext, hertz, phase, amp, 2), 2, 48000).describe()['launches']]
    finally:
        _set_default('fuse_reduce', 1)
    assert kinds == ['chain', 'reduce']
    assert _lib.lib().sigb_set_default_option(b'nope', 1) == _lib.SIGB_EINVAL


def test_pansum_shape_errors(ns, engine):
    prm = cases.instance_params(5, 40)
    with pytest.raises(chain.BadShape):
        engine.compile(cases.build_instances(ns, ext, prm), 3, 48000)      # PanSum yields 2 channels


def test_mix_of_oscillators_is_one_launch(ns, engine):
    """`Gain and mix nodes fuse into their producers`: Mix(Sine, Sawtooth) and RingMod(Sine, Triangle) are ONE
    chain launch each (the second oscillator stays in registers); a shared operand (fan-out) is not fused."""
    d = engine.compile(cases.CASES_BY_NAME['mix'].build(ns), 3, 48000).describe()
    (l,) = d['launches']
    assert l['kind'] == 'chain' and l['epilogue'] == 'mix' and l['other'] == 'osc' and d['buffers'] == 0
    d = engine.compile(cases.CASES_BY_NAME['ringmod'].build(ns), 2, 48000).describe()
    (l,) = d['launches']
    assert l['epilogue'] == 'ringmod' and l['other'] == 'osc'
    kinds = [x['kind'] for x in engine.compile(cases.CASES_BY_NAME['fanout'].build(ns), 4, 48000).describe()['launches']]
    assert 'ewise' in kinds
    _set_default('fuse_pointwise', 0)
    try:
        kinds = [x['kind'] for x in engine.compile(cases.CASES_BY_NAME['mix'].build(ns), 3, 48000).describe()['launches']]
    finally:
        _set_default('fuse_pointwise', 1)
    assert kinds == ['chain', 'chain', 'ewise']


def test_modulated_pan_lowers_to_a_parameter_row(ns, engine):
    """PanSum.pan driven by an emitter (an LFO sweeping the stereo position): a block-rate port like the others, sampled once
    per request on the device.  The voices stay fused (k_pan_weights re-derives their (L, R) weights per request); without the
    fusion k_reduce reads the pan row."""
    prm = cases.instance_params(5, 64)
    ps = cases.build_instances(ns, ext, prm)
    assert [l['kind'] for l in engine.compile(ps, 2, 48000).describe()['launches']] == ['voices']
    ps.pan = cases.sweep(ns, [np.full(64, 0.1)], [np.full(64, 0.9)], [[0.3]], [[0.0]])
    d = engine.compile(ps, 2, 48000).describe()
    assert [l['kind'] for l in d['launches']] == ['voices'] and d['modulated_parameters'] >= 1
    _lib.lib().sigb_set_default_option(b'fuse_reduce', 0)
    try:
        kinds = [l['kind'] for l in engine.compile(ps, 2, 48000).describe()['launches']]
    finally:
        _lib.lib().sigb_set_default_option(b'fuse_reduce', 1)
    assert 'voices' not in kinds and kinds[-1] == 'reduce'
