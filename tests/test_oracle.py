"""CPU: pins the oracle (oracle/np_oracle.py) against golden vectors minted from the UNMODIFIED
reference (oracle/make_golden.py), and the restated third-party algorithms against scipy."""
import json
import os

import numpy as np
import pytest
import scipy.signal

from conftest import GOLDEN_DIR, load_golden, max_abs_err
from oracle import cases, np_oracle, ref_harness


@pytest.mark.parametrize('case', cases.CASES, ids=lambda c: c.name)
def test_oracle_matches_reference_golden(case, ns):
    """GraphOracle walking the signals_b200.chain mirror == the reference's own render, bit for bit
    (same numpy/scipy calls in the same order), for every case the reference was run on."""
    if case.frames > 500000 and os.environ.get('SIGB_FAST_TESTS'):
        pytest.skip('long render')
    want = load_golden(case.name)
    graph, orc = case.build(ns), np_oracle.GraphOracle(case.rate)
    if case.block:
        got = np.concatenate([orc.render(graph, case.position + r, min(case.block, case.frames - r), case.channels)
                              for r in range(0, case.frames, case.block)])
    else:
        got = orc.render(graph, case.position, case.frames, case.channels)
    got = got[::case.stride]
    assert got.shape == want.shape
    assert max_abs_err(got, want) == 0.0


def test_golden_meta_lists_every_case():
    with open(os.path.join(GOLDEN_DIR, 'META.json')) as f:
        meta = json.load(f)
    assert set(meta['cases']) == {c.name for c in cases.CASES}
    assert meta['reference_pins'] == {'numpy': '1.23.0', 'scipy': '1.10.1'}
    assert set(meta['error_cases']) == {e[0] for e in cases.ERROR_CASES + cases.RUNTIME_ERROR_CASES}


@pytest.mark.parametrize('name,build,frames,channels,exc', cases.ERROR_CASES + cases.RUNTIME_ERROR_CASES, ids=lambda v: v if isinstance(v, str) else '')
def test_oracle_error_cases(name, build, frames, channels, exc, ns):
    with pytest.raises(Exception) as info:
        np_oracle.GraphOracle().render(build(ns), 0, frames, channels)
    assert any(k.__name__ == exc for k in type(info.value).__mro__)


@pytest.mark.skipif(not ref_harness.available(), reason='reference sources only exist in the build container')
@pytest.mark.parametrize('case', [c for c in cases.CASES if c.frames <= 48000], ids=lambda c: c.name)
def test_oracle_walks_reference_objects(case):
    """The same GraphOracle evaluates graphs built from the reference's OWN node classes and matches
    the reference's recursion on them (checks the oracle, not the mirror)."""
    ref = ref_harness.load()
    rns = cases.ref_namespace(ref)
    want = ref_harness.render(ref, case.build(rns), case.position, case.frames, case.channels, case.rate)
    got = np_oracle.GraphOracle(case.rate).render(case.build(rns), case.position, case.frames, case.channels)
    assert max_abs_err(got, np.broadcast_to(want, got.shape)) == 0.0


@pytest.mark.parametrize('btype', ['lp', 'hp'])
@pytest.mark.parametrize('wn', [100 / 24000, 0.025, 0.3, 0.9])
def test_closed_form_butterworth_matches_scipy(btype, wn):
    want = scipy.signal.butter(2, wn, btype, output='sos')
    got = np_oracle.butter2_closed_form(btype, wn)
    assert np.abs(got - want).max() < 1e-14


def test_df2t_restatement_matches_sosfilt():
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (2000, 3))
    sos = scipy.signal.butter(6, 0.1, 'lp', output='sos')
    want = scipy.signal.sosfilt(sos, x, axis=0)
    got, _ = np_oracle.sosfilt_df2t(sos, x)
    assert np.abs(got - want).max() < 1e-12


def test_example_sine_formula_vs_graph_form():
    """scripts/example_sine.py:50-53 vs the Sine->Gain graph (config C1): op order differs, values agree to 1e-12."""
    script = np_oracle.example_sine_block(0, 4800, 48000.0)
    graph = load_golden('example_sine')
    assert np.abs(script - graph).max() < 1e-12


def test_blockwise_oscillators_equal_single_request(ns):
    case = cases.CASES_BY_NAME['square_edges']
    g = case.build(ns)
    o = np_oracle.GraphOracle()
    whole = o.render(g, 0, 4800, case.channels)
    parts = np.concatenate([o.render(g, p, n, case.channels) for p, n in ((0, 1), (1, 999), (1000, 3000), (4000, 800))])
    assert np.array_equal(whole, parts)


def test_render_voice_chain_equals_graph_oracle(ns):
    hertz, phase, cutoff, g = cases.voice_params(2, 8)
    arr = np_oracle.render_voice_chain(0, 4800, 48000, hertz, phase, cutoff, g)
    graph = load_golden('lowpass_c2_8v')[:4800]
    assert np.abs(arr - graph).max() < 1e-15
