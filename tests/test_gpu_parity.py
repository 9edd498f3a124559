"""GPU: parity of the CUDA path (through the C ABI) against golden vectors minted from the
reference, against the numpy oracle on seeded inputs, and size-independent properties at the
BASELINE sizes.  Tolerances are the north-star's: 1e-6 max-abs for oscillators / gain / mix,
1e-4 for cascaded IIR output (each case carries its own in oracle/cases.py)."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, max_abs_err
from oracle import cases, np_oracle

pytestmark = pytest.mark.gpu

RATE = 48000


def _torch():
    import torch
    return torch


def render_case(engine, ns, case, **options):
    compiled = engine.compile(case.build(ns), case.channels, case.rate, case.frames)
    for k, v in {**case.options, **options}.items():
        compiled.set_option(k, v)
    if case.block:       # consecutive requests: block-rate parameters are re-sampled at each request's first frame
        out = np.concatenate([compiled.render_device(case.position + r, min(case.block, case.frames - r)).cpu().numpy()
                              for r in range(0, case.frames, case.block)])
    else:
        out = compiled.render_device(case.position, case.frames).cpu().numpy()
    launches = compiled.launch_count
    compiled.close()
    assert launches > 0, 'no CUDA kernel was launched'
    return out


FILTER_CASES = [c for c in cases.CASES if c.tol > 1e-6]
PLAIN_CASES = [c for c in cases.CASES if c.tol <= 2e-6]


@pytest.mark.parametrize('case', cases.CASES, ids=lambda c: c.name)
def test_cuda_matches_reference_golden(case, ns, engine):
    got = render_case(engine, ns, case)[::case.stride]
    want = load_golden(case.name)
    assert got.shape == want.shape and got.dtype == np.float32
    if case.name == 'amp_frac':
        # x ** 0.5 is NaN for x < 0: at the sine's zero crossings the sign of a ~1e-17 residue decides
        # NaN vs 0 in the reference itself, and d/dx x**0.5 blows up near 0: keep the well-conditioned |x| > 0.25
        a = load_golden('amp_int')          # same oscillator through exponents [2, 3]: recover |x|
        keep = np.stack([np.abs(a[:, 0]) ** 0.5, np.abs(a[:, 1]) ** (1 / 3)], axis=1) > 0.25
        got, want = np.where(keep, got, 0.0), np.where(keep, want, 0.0)
    err = max_abs_err(got, want)
    assert err <= case.tol, f'{case.name}: max-abs {err:.3e} > {case.tol:.1e}'


@pytest.mark.parametrize('case', FILTER_CASES, ids=lambda c: c.name)
@pytest.mark.parametrize('mode', ['seq', 'scan9', 'scan16', 'scan18', 'pipe', 'pipe2', 'oscreg', 'oscsvf'])
def test_every_filter_kernel_variant_matches_golden(case, mode, ns, engine):
    """k_chain_seq, the time-parallel kernels (scan9 = k_chain_scan2, deep cascades forced onto it too; scan16 / scan18 =
    k_chain_scan3 with the float64 / packed float32 carry chain), the section-pipelined k_cascade_pipe (forced from 2
    sections) and the register-resident oscillator kernels (from one section: 'oscreg' = k_osc_delta for second-order
    sections from two on, k_osc_reg otherwise; 'oscsvf' = k_osc_reg's state-variable sections throughout) agree with the
    reference."""
    opts = (dict(force_seq=1) if mode == 'seq' else dict(cascade_pipe=1, osc_reg=0) if mode == 'pipe' else dict(cascade_pipe=1, pipe_spw=2, osc_reg=0) if mode == 'pipe2'
            else dict(osc_reg=1) if mode == 'oscreg' else dict(osc_reg=1, osc_delta=0) if mode == 'oscsvf'
            else dict(scan_variant=int(mode[4:]), cascade_pipe=0))
    got = render_case(engine, ns, case, **opts)[::case.stride]
    err = max_abs_err(got, load_golden(case.name))
    assert err <= case.tol, f'{case.name}/{mode}: max-abs {err:.3e}'


def test_sine_error_budget(ns, engine):
    """Oscillator budget: <= 1e-6 full scale; report how much of it each sin2pi variant uses."""
    torch = _torch()
    from signals_b200 import _lib
    lib = _lib.lib()
    r = np.concatenate([np.linspace(-0.5, 0.5, 2_000_001), np.random.default_rng(0).uniform(-0.5, 0.5, 1_000_000)])
    want = np.sin(2 * np.pi * r.astype(np.float32).astype(np.float64))
    d_r = torch.from_numpy(r).cuda()
    d_o = torch.empty(r.size, dtype=torch.float32, device='cuda')
    errs = {}
    for variant in (0, 1, 2):
        st = lib.sigb_probe_sin(ctypes.c_void_p(d_r.data_ptr()), r.size, ctypes.c_void_p(d_o.data_ptr()), variant, None)
        assert st == 0
        torch.cuda.synchronize()
        errs[variant] = float(np.abs(d_o.cpu().numpy().astype(np.float64) - want).max())
    print('sin2pi max-abs error by variant (0=MUFU, 1=folded MUFU, 2=folded polynomial):', errs)
    assert errs[1] < 5e-7 and errs[2] < 3e-7


def test_oscillators_are_block_invariant_bitwise(ns, engine):
    """Rendering [0,F) in one call equals rendering it in unequal calls, bit for bit."""
    for wave in ('Sine', 'Square', 'Sawtooth', 'Triangle'):
        g = cases.osc(ns, wave, [[440.0, 439.99, 12000.0, 27.5]], [[0.0, 0.1, 0.2, 0.3]])
        compiled = engine.compile(g, 4, RATE)
        whole = compiled.render_device(0, 10000).cpu().numpy()
        cuts = [0, 1, 48, 1000, 1001, 4097, 10000]
        parts = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
        assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32)), wave
        compiled.close()


def test_filter_state_is_carried_across_calls(ns, engine):
    """Streaming a filter chain in ragged chunks == one request (the reference cannot do this:
    it restarts from zero state every block)."""
    case = cases.CASES_BY_NAME['lowpass_c2_8v']
    want = load_golden(case.name)
    compiled = engine.compile(case.build(ns), case.channels, RATE)
    cuts = [0, 1, 7, 512, 513, 5000, 20000, 20112, 48000]
    parts = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
    assert max_abs_err(parts, want) <= 1e-5
    compiled.close()


def test_seek_restarts_with_context_like_the_reference(ns, engine):
    """After a seek the plan zeroes state and warms up on context_frames() frames, which is
    exactly what the reference does for every block (fx.py:93-105)."""
    case = cases.CASES_BY_NAME['lowpass_blockwise']
    compiled = engine.compile(case.build(ns), case.channels, RATE)
    compiled.render_device(0, 1000)                       # some unrelated history
    got = compiled.render_device(case.position, case.frames).cpu().numpy()
    assert max_abs_err(got, load_golden(case.name)) <= 1e-5
    compiled.close()


def test_render_host_equals_render_device(ns, engine):
    case = cases.CASES_BY_NAME['lowpass_c2_8v']
    compiled = engine.compile(case.build(ns), case.channels, RATE)
    compiled.set_option('host_slab_bytes', 8 * 4 * 1000)   # force many slabs
    dev = compiled.render_device(0, case.frames).cpu().numpy()
    host = compiled.render_host(0, case.frames)
    assert max_abs_err(host, dev) <= 2e-6
    assert max_abs_err(host, load_golden(case.name)) <= 1e-5
    compiled.close()


@pytest.mark.parametrize('voices,frames,wave,cls', [
    (64, 48000, 'Sine', 'LowPass'), (200, 9999, 'Sawtooth', 'HighPass'), (33, 12345, 'Triangle', 'LowPass'),
    (1, 4800, 'Square', 'LowPass'), (1000, 4999, 'Sine', 'LowPass')])
def test_voice_chain_vs_oracle(voices, frames, wave, cls, ns, engine):
    """Config C2 shape on seeded inputs, ragged sizes (voices not a multiple of 32, frames not a
    multiple of the scan step)."""
    hertz, phase, cutoff, g = cases.voice_params(1234 + voices, voices)
    graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, wave, [hertz], [phase]), [cutoff], cls), [g])
    want = np_oracle.render_voice_chain(0, frames, RATE, hertz, phase, cutoff, g, wave=wave,
                                        btype=np_oracle.FILTER_TYPES[cls])
    compiled = engine.compile(graph, voices, RATE)
    got = compiled.render_device(0, frames).cpu().numpy()
    compiled.close()
    err = max_abs_err(got, want)
    assert err <= 1e-4, err
    print(f'C2-shape {voices}x{frames} {wave}->{cls}: max-abs {err:.3e}')


def test_low_cutoff_iir_stays_inside_budget(ns, engine):
    """Direct-form float32 biquads lose 1e-3 at 20-100 Hz cutoffs; the state-variable sections must not."""
    v = 16
    hertz = np.linspace(20.0, 200.0, v)
    cutoff = np.geomspace(20.0, 300.0, v)
    graph = cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz]), [cutoff])
    frames = 10 * RATE
    want = np_oracle.render_voice_chain(0, frames, RATE, hertz, np.zeros(v), cutoff, np.ones(v))
    for opts in (dict(force_seq=1), dict(scan_variant=9), dict(scan_variant=18)):
        compiled = engine.compile(graph, v, RATE)
        for k, val in opts.items():
            compiled.set_option(k, val)
        got = compiled.render_device(0, frames).cpu().numpy()
        compiled.close()
        err = max_abs_err(got, want)
        assert err <= 2e-5, (opts, err)


def test_group_sum_and_pan_sum_vs_numpy(ns, engine):
    from signals_b200.chain import ext
    rng = np.random.default_rng(3)
    p, groups, frames = 4096, 8, 4800
    hertz = rng.uniform(27.5, 12000.0, p)
    phase = rng.uniform(0, 1, p)
    amp = rng.uniform(0, 1, p) / (p // groups)
    bank = cases.gain(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [amp])
    gs = ext.GroupSum()
    gs.get_state().groups = groups
    gs.input = bank
    x = np_oracle.gain(np_oracle.sine(np_oracle.osc_cycles(0, frames, RATE, hertz[None], phase[None])), amp[None])
    got = engine.compile(gs, groups, RATE).render_device(0, frames).cpu().numpy()
    assert max_abs_err(got, np_oracle.group_sum(x, groups)) <= 1e-6
    ps = ext.PanSum()
    ps.input = bank
    pan = rng.uniform(0, 1, p)
    ps.pan = cases.fixed(ns, [pan])
    got = engine.compile(ps, 2, RATE).render_device(0, frames).cpu().numpy()
    assert max_abs_err(got, np_oracle.pan_sum(x, pan[None])) <= 1e-6


def test_buffer_source_through_cascade(ns, engine):
    """Config C4 shape: an HBM-resident noise block through 8 chained LowPass nodes, streamed in
    chunks with carried state, vs scipy's sosfilt cascade."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(4)
    ch, frames = 96, 20000
    x = rng.uniform(-1, 1, (frames, ch)).astype(np.float32)
    cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, ch)))
    node = ext.Buffer(x)
    for s in range(8):
        node = cases.lowpass(ns, node, [cut[s]])
    want, _ = np_oracle.render_cascade(x.astype(np.float64), cut, RATE)
    compiled = engine.compile(node, ch, RATE)
    cuts = [0, 4800, 4801, 12000, 20000]
    got = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
    past_end = compiled.render_device(frames, 64).cpu().numpy()
    compiled.close()
    assert max_abs_err(got, want) <= 1e-4
    assert np.isfinite(past_end).all()


def test_full_size_voice_bank_properties(ns, engine):
    """BASELINE config C2 at full size (4096 voices x 10 s, 7.9 GB): spot-check 12 voices against
    the oracle over all 480000 frames, exact gain linearity, and finiteness of everything."""
    torch = _torch()
    v, frames = 4096, 10 * RATE
    hertz, phase, cutoff, g = cases.voice_params(2, v)
    graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])
    compiled = engine.compile(graph, v, RATE)
    out = compiled.render_device(0, frames)
    assert bool(torch.isfinite(out).all())
    pick = np.concatenate([np.argsort(cutoff)[:4], np.argsort(hertz)[:4], np.random.default_rng(0).choice(v, 4)])
    got = out[:, torch.from_numpy(pick).cuda()].cpu().numpy()
    want = np_oracle.render_voice_chain(0, frames, RATE, hertz[pick], phase[pick], cutoff[pick], g[pick])
    err = max_abs_err(got, want)
    print(f'C2 full size: max-abs over 12 voices x 480000 frames = {err:.3e}')
    assert err <= 1e-4
    compiled.close()
    graph2 = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [2.0 * g])
    compiled2 = engine.compile(graph2, v, RATE)
    out2 = compiled2.render_device(0, frames)
    assert bool(torch.equal(out2, out * 2.0))          # power-of-two gain is exact in float32
    compiled2.close()


def test_time_split_pieces_match_oracle(ns, engine):
    """20 tiles on 148 SMs: the packed kernel cuts every tile along time into pieces that warm the
    filters up from zero state (decayed below 2^-40); voices in every piece must still match the oracle,
    and the state handed to the next call must be the true one."""
    torch = _torch()
    v, frames = 640, 10 * RATE
    hertz, phase, cutoff, g = cases.voice_params(77, v)
    cutoff[::7] = 100.0                                  # slowest-decaying filters the config allows
    graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])
    compiled = engine.compile(graph, v, RATE)
    first = compiled.render_device(0, frames - 4800)
    second = compiled.render_device(frames - 4800, 4800)              # continues from the carried state
    pick = np.unique(np.concatenate([np.arange(0, v, 7)[:12], np.random.default_rng(1).choice(v, 12)]))
    idx = torch.from_numpy(pick).cuda()
    got = torch.cat([first[:, idx], second[:, idx]]).cpu().numpy()
    want = np_oracle.render_voice_chain(0, frames, RATE, hertz[pick], phase[pick], cutoff[pick], g[pick])
    err = max_abs_err(got, want)
    print(f'time-split pieces: max-abs over {len(pick)} voices x {frames} frames = {err:.3e}')
    assert err <= 1e-4
    compiled.set_option('scan_split', 0)
    whole = compiled.render_device(0, frames - 4800)
    compiled.set_option('scan_split', 1)
    assert float((whole - first).abs().max()) <= 2e-6
    compiled.close()


# ------------------------------------------------------------------------------------------------
# fused render + mix-down kernels (configs C3 and C5)
# ------------------------------------------------------------------------------------------------

def _set_default(key, value):
    from signals_b200 import _lib
    assert _lib.lib().sigb_set_default_option(key.encode(), int(value)) == 0


@pytest.mark.parametrize('partials,groups,frames,position', [
    (4096, 8, 4800, 0),            # C3 in miniature: 512 partials per group
    (4096, 4, 5000, 12345),        # ragged last tile, 1024 per group, position > 0
    (300, 3, 777, 2 ** 31 + 5),    # groups not a multiple of 8, 100 per group, huge position
    (9000, 2, 600, 47999),         # 4500 per group: more than one shared-memory chunk
])
def test_bank_fused_vs_oracle(partials, groups, frames, position, ns, engine):
    """k_bank (oscillator bank fused with GroupSum) against the float64 numpy oracle."""
    from signals_b200.chain import ext
    hertz, phase, amp = cases.bank_params(3, partials, partials // groups)
    compiled = engine.compile(cases.build_bank(ns, ext, hertz, phase, amp, groups), groups, RATE)
    assert [l['kind'] for l in compiled.describe()['launches']] == ['bank']
    got = compiled.render_device(position, frames).cpu().numpy()
    compiled.close()
    want = np_oracle.render_bank(position, frames, RATE, hertz, phase, amp, groups)
    err = max_abs_err(got, want)
    print(f'bank {partials}->{groups}: max-abs {err:.3e}')
    # at position 2^31 the REFERENCE's own float64 phase (n / rate * hertz ~ 5.4e8 cycles at 12 kHz) is rounded to 1.2e-7
    # cycles = 7.5e-7 rad per partial, while the Q0.64 phase here is exact: the 2e-6 there is the reference's rounding
    assert err <= (2e-6 if position > 2 ** 31 else 1e-6)


def test_bank_full_scale_partials_stay_inside_budget(ns, engine):
    """Two partials of amplitude 0.5 per group: the per-partial error budget (phase-word drift + MUFU)
    must itself stay under 1e-6 of full scale, not only its average over 1024 small partials."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(33)
    hertz = rng.uniform(27.5, 12000.0, 8)
    phase = rng.uniform(0, 1, 8)
    amp = np.full(8, 0.5)
    compiled = engine.compile(cases.build_bank(ns, ext, hertz, phase, amp, 4), 4, RATE)
    got = compiled.render_device(0, 96000).cpu().numpy()
    compiled.close()
    assert max_abs_err(got, np_oracle.render_bank(0, 96000, RATE, hertz, phase, amp, 4)) <= 1e-6


def test_bank_fused_equals_materialised_path(ns, engine):
    from signals_b200.chain import ext
    hertz, phase, amp = cases.bank_params(31, 2048, 256)
    graph = lambda: cases.build_bank(ns, ext, hertz, phase, amp, 8)   # noqa: E731
    fused = engine.compile(graph(), 8, RATE)
    a = fused.render_device(100, 3000).cpu().numpy()
    fused.close()
    _set_default('fuse_reduce', 0)
    try:
        plain = engine.compile(graph(), 8, RATE)
        assert [l['kind'] for l in plain.describe()['launches']] == ['chain', 'reduce']
        b = plain.render_device(100, 3000).cpu().numpy()
        plain.close()
    finally:
        _set_default('fuse_reduce', 1)
    assert max_abs_err(a, b) <= 5e-7


@pytest.mark.parametrize('m', [1, 4])
def test_instances_fused_vs_oracle(m, ns, engine):
    """k_voices (config C5 in miniature): 3000 randomised osc/filter/gain/pan instances -> stereo."""
    from signals_b200.chain import ext
    prm = cases.instance_params(5, 3000)
    frames = 4800
    _set_default('voices_m', m)
    try:
        compiled = engine.compile(cases.build_instances(ns, ext, prm), 2, RATE)
    finally:
        _set_default('voices_m', 0)
    d = compiled.describe()
    assert [l['kind'] for l in d['launches']] == ['voices'] and d['launches'][0]['channels_per_thread'] == m
    cuts = [0, 1000, 1001, 1017, frames]                       # ragged calls: state and phase are carried
    got = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
    compiled.close()
    want = np_oracle.render_instances(prm, 0, frames, RATE)
    err = max_abs_err(got, want)
    print(f'instances M={m}: max-abs {err:.3e}, mix peak {np.abs(want).max():.3f}')
    assert err <= 1e-6


def test_instances_discontinuities_land_on_the_reference_sample(ns, engine):
    """Rational hertz/rate puts samples exactly ON the Square/Sawtooth jumps: the phase-word
    fast path must hand those tiles to the float64 path, or single samples are off by 1-2 full scale."""
    from signals_b200.chain import ext
    hz = np.array([440.0, 1000.0, 12000.0, 6000.0, 439.99, 100.0])
    n = hz.size
    for wave in (1, 2, 3):
        prm = dict(wave=np.full(n, wave), filt=np.zeros(n, dtype=int), hertz=hz, phase=np.array([0, 0, 0, 0, 0.3, -0.75]),
                   cutoff=np.full(n, 1000.0), gain=np.full(n, 1.0 / n), pan=np.linspace(0.1, 0.9, n))
        compiled = engine.compile(cases.build_instances(ns, ext, prm), 2, RATE)
        got = compiled.render_device(0, 48000).cpu().numpy()
        compiled.close()
        want = np_oracle.render_instances(prm, 0, 48000, RATE)
        assert max_abs_err(got, want) <= 1e-6, wave


def test_instances_seek_and_materialised_path_agree(ns, engine):
    """position > 0 on a fresh plan: zero state + context warm-up (fx.py:93-105), same as the graph
    oracle; and the fused kernel equals the chain + merge + reduce launches it replaces."""
    from signals_b200.chain import ext
    prm = cases.instance_params(55, 96)
    graph = lambda: cases.build_instances(ns, ext, prm)   # noqa: E731
    compiled = engine.compile(graph(), 2, RATE)
    got = compiled.render_device(4800, 512).cpu().numpy()
    compiled.close()
    want = np_oracle.GraphOracle(RATE).render(graph(), 4800, 512, 2)
    assert max_abs_err(got, want) <= 1e-6
    _set_default('fuse_reduce', 0)
    try:
        plain = engine.compile(graph(), 2, RATE)
        assert 'voices' not in [l['kind'] for l in plain.describe()['launches']]
        b = plain.render_device(4800, 512).cpu().numpy()
        plain.close()
    finally:
        _set_default('fuse_reduce', 1)
    assert max_abs_err(got, b) <= 1e-6


def test_cascade_pipe_ragged_channels_and_streaming(ns, engine):
    """k_cascade_pipe: channel counts that are not a multiple of the 64-channel tile (odd, so the last
    lane holds one live channel), rows that are not a multiple of the 16-row chunk, state carried across
    ragged calls, against scipy's float64 sosfilt cascade; and bit-agreement of a streamed render with the
    single-request render."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(44)
    for ch, nsec in [(200, 8), (68, 5), (67, 4), (4, 3)]:
        frames = 6000
        x = rng.uniform(-1, 1, (frames, ch)).astype(np.float32)
        cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (nsec, ch)))
        node = ext.Buffer(x)
        for s in range(nsec):
            node = cases.lowpass(ns, node, [cut[s]], 'HighPass' if s == 1 else 'LowPass')
        want = x.astype(np.float64)
        for s in range(nsec):
            want, _ = np_oracle.render_cascade(want, cut[s:s + 1], RATE, btype='hp' if s == 1 else 'lp')
        compiled = engine.compile(node, ch, RATE)
        whole = compiled.render_device(0, frames).cpu().numpy()
        compiled.reset()
        cuts = [0, 1, 17, 1000, 1016, 4803, frames]
        parts = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
        compiled.close()
        err = max_abs_err(whole, want)
        print(f'cascade pipe {ch} ch x {nsec} sections: max-abs {err:.3e}')
        assert err <= 1e-4
        assert np.array_equal(parts, whole)      # sequential in time: chunking cannot change a single bit
    # the same with cascades of ONE kind on a materialised block: k_cascade_reg (all sections in registers;
    # ragged channel counts take its guarded path, multiples of 64 its 8-byte path, both block widths)
    # variant 0 on whole 64-channel tiles of low-pass sections = k_cascade_delta (five operations per section); 4 keeps the
    # state-variable sections of k_cascade_reg in 8-row blocks
    for ch, nsec, btype, variant in [(200, 8, 'lp', 0), (68, 5, 'hp', 0), (67, 4, 'lp', 0), (4, 3, 'hp', 0), (128, 8, 'lp', 0),
                                     (128, 8, 'lp', 1), (64, 3, 'hp', 1), (320, 6, 'lp', 0), (64, 7, 'hp', 0), (128, 8, 'lp', 4),
                                     (64, 3, 'lp', 0), (192, 7, 'lp', 0), (64, 4, 'lp', 0), (64, 5, 'lp', 0), (128, 8, 'hp', 0), (64, 4, 'hp', 0),
                                     (64, 6, 'hp', 0), (64, 5, 'hp', 4)]:
        frames = 6000
        x = rng.uniform(-1, 1, (frames, ch)).astype(np.float32)
        cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (nsec, ch)))
        node = ext.Buffer(x)
        for s in range(nsec):
            node = cases.lowpass(ns, node, [cut[s]], 'HighPass' if btype == 'hp' else 'LowPass')
        want = x.astype(np.float64)
        for s in range(nsec):
            want, _ = np_oracle.render_cascade(want, cut[s:s + 1], RATE, btype=btype)
        compiled = engine.compile(node, ch, RATE)
        compiled.set_option('reg_variant', variant)
        pieces = compiled.render_device(0, frames).cpu().numpy()       # short warm-ups: cut into pieces along time
        compiled.set_option('pipe_segments', 1)
        compiled.reset()
        whole = compiled.render_device(0, frames).cpu().numpy()
        compiled.reset()
        cuts = [0, 1, 17, 1000, 1016, 4803, frames]
        parts = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
        assert max_abs_err(pieces, whole) <= 1e-6
        compiled.set_option('cascade_reg', 0)           # the section-pipelined kernel on the same plan and state convention
        compiled.reset()
        piped = compiled.render_device(0, frames).cpu().numpy()
        compiled.close()
        err = max_abs_err(whole, want)
        print(f'cascade reg {ch} ch x {nsec} {btype} sections (variant {variant}): max-abs {err:.3e}, vs k_cascade_pipe {max_abs_err(whole, piped):.3e}')
        assert err <= 1e-4
        assert max_abs_err(whole, piped) <= 2e-5
        if btype == 'lp' and ch % 64 == 0 and variant == 0:
            # delta form, low-pass: the recurrent states survive the hand-over bit for bit, the one-row memory of a section's
            # second zero is re-derived from them (one float32 rounding): chunked and unchunked differ by rounding noise
            # (high-pass sections in delta form have no such memory and stay bit-exact)
            assert max_abs_err(parts, whole) <= 5e-7
        else:
            assert np.array_equal(parts, whole)      # sequential in time: chunking cannot change a single bit
    # odd channel count (one live channel in the last lane) from an oscillator source, into a padded block
    torch = _torch()
    ch = 67
    hertz = rng.uniform(55.0, 880.0, ch)
    node = cases.osc(ns, 'Sawtooth', [hertz], [rng.uniform(0, 1, ch)])
    cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (3, ch)))
    for s in range(3):
        node = cases.lowpass(ns, node, [cut[s]])
    compiled = engine.compile(node, ch, RATE)
    assert compiled.describe()['launches'][0]['sections'] == 3
    big = torch.full((3000, 68), 7.0, dtype=torch.float32, device='cuda')
    compiled.render_device(0, 3000, big[:, :ch])
    compiled.close()
    got = big.cpu().numpy()
    assert (got[:, ch:] == 7.0).all()            # the padding column is untouched
    want = np_oracle.GraphOracle(RATE).render(node, 0, 3000, ch)
    assert max_abs_err(got[:, :ch], want) <= 1e-4


@pytest.mark.parametrize('kernel', ['pipe', 'reg', 'reg_hp', 'reg_svf', 'reg_svf_hp', 'reg_r4', 'reg_ragged', 'reg_4sec', 'reg_ragged_rows'])
def test_cascade_pipe_time_segments_match_oracle(kernel, ns, engine):
    """k_cascade_pipe / k_cascade_reg cut long renders into time segments that warm up from zero state
    (decayed below 2^-40); every segment must match the float64 cascade, the segmented render must agree
    with the unsegmented one, and the state handed to the next call must be the true one."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(45)
    ch, nsec, frames = (190 if kernel == 'reg_ragged' else 192), (4 if kernel == 'reg_4sec' else 8), (60003 if kernel == 'reg_ragged_rows' else 60000)
    cls, btype = ('HighPass', 'hp') if kernel in ('reg_svf_hp', 'reg_hp') else ('LowPass', 'lp')
    x = rng.uniform(-1, 1, (frames + 4000, ch)).astype(np.float32)
    cut = np.exp(rng.uniform(np.log(600.0), np.log(8000.0), (nsec, ch)))
    node = ext.Buffer(x)
    for s in range(nsec):
        node = cases.lowpass(ns, node, [cut[s]], cls)
    compiled = engine.compile(node, ch, RATE)
    compiled.set_option('cascade_reg', 0 if kernel == 'pipe' else -1)
    # 0: k_cascade_delta on this all-low-pass cascade ('reg'; 'reg_ragged' has a ragged last tile and stays on k_cascade_reg),
    # 4: k_cascade_reg's state-variable sections in 8-row blocks
    compiled.set_option('reg_variant', 1 if kernel == 'reg_r4' else 4 if kernel.startswith('reg_svf') else 0)
    warm = compiled.describe()['launches'][0]['warm_rows']
    assert 0 < warm < frames // 8, warm                       # so that the launch really is segmented
    first = compiled.render_device(0, frames).cpu().numpy()
    second = compiled.render_device(frames, 4000).cpu().numpy()          # continues from the carried state
    compiled.set_option('pipe_segments', 1)
    compiled.reset()
    whole = compiled.render_device(0, frames).cpu().numpy()
    compiled.close()
    pick = rng.choice(ch, 24, replace=False)
    want, _ = np_oracle.render_cascade(x[:, pick].astype(np.float64), cut[:, pick], RATE, btype)
    err = max_abs_err(np.concatenate([first, second])[:, pick], want)
    print(f'{kernel} segments: warm_rows {warm}, max-abs over 24 channels x {frames + 4000} frames = {err:.3e}')
    assert err <= 1e-4
    assert max_abs_err(first, whole) <= 1e-6


# ------------------------------------------------------------------------------------------------
# BASELINE full sizes of C3 / C4 / C5: size-independent properties + spot checks against the oracle
# ------------------------------------------------------------------------------------------------

def test_full_size_bank_properties(ns, engine):
    """Config C3 at full size (65,536 partials -> 64 channels x 10 s): windows at the start, middle and end
    of the render against the float64 oracle for four groups, and exact linearity in the amplitudes."""
    from signals_b200.chain import ext
    torch = _torch()
    p, groups, frames = 65536, 64, 10 * RATE
    hertz, phase, amp = cases.bank_params(3, p, p // groups)
    compiled = engine.compile(cases.build_bank(ns, ext, hertz, phase, amp, groups), groups, RATE)
    out = compiled.render_device(0, frames)
    compiled.close()
    assert bool(torch.isfinite(out).all())
    per = p // groups
    for g in (0, 17, 40, 63):
        sl = slice(g * per, (g + 1) * per)
        for pos in (0, 240000 - 128, frames - 1000):
            want = np_oracle.render_bank(pos, 1000, RATE, hertz[sl], phase[sl], amp[sl], 1)
            got = out[pos:pos + 1000, g].cpu().numpy()[:, None]
            assert max_abs_err(got, want) <= 1e-6, (g, pos)
    doubled = engine.compile(cases.build_bank(ns, ext, hertz, phase, 2.0 * amp, groups), groups, RATE)
    out2 = doubled.render_device(0, frames)
    doubled.close()
    assert bool(torch.equal(out2, out * 2.0))          # power-of-two amplitude scaling is exact in float32


def test_full_size_cascade_properties(ns, engine):
    """Config C4 at full size (16,384 channels x 8 low-pass sections x 60 s, streamed in 5 s slabs with carried
    state, as bench.py --config c4 does): six channels over the whole minute against scipy's float64 cascade."""
    from signals_b200.chain import ext
    torch = _torch()
    ch, slab, n_slabs = 16384, 5 * RATE, 12
    rng = np.random.default_rng(4)
    cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, ch)))
    gen = torch.Generator(device='cuda')
    gen.manual_seed(4)
    noise = torch.rand((slab, ch), generator=gen, device='cuda', dtype=torch.float32) * 2 - 1
    buf = ext.Buffer(noise)
    node = buf
    for s in range(8):
        node = cases.lowpass(ns, node, [cut[s]])
    compiled = engine.compile(node, ch, RATE, slab)
    assert [l['kind'] for l in compiled.describe()['launches']] == ['chain']
    pick = np.concatenate([np.argsort(cut.min(0))[:3], rng.choice(ch, 3, replace=False)])
    idx = torch.from_numpy(pick).cuda()
    out = torch.empty((slab, ch), dtype=torch.float32, device='cuda')
    got = []
    for k in range(n_slabs):
        compiled.bind_window(buf, noise, k * slab)
        compiled.render_device(k * slab, slab, out)
        got.append(out[:, idx].cpu().numpy())
    assert bool(torch.isfinite(out).all())
    compiled.close()
    got = np.concatenate(got)
    x = np.tile(noise[:, idx].cpu().numpy().astype(np.float64), (n_slabs, 1))
    want, _ = np_oracle.render_cascade(x, cut[:, pick], RATE)
    err = max_abs_err(got, want)
    print(f'C4 full size: max-abs over 6 channels x {n_slabs * slab} frames = {err:.3e}')
    assert err <= 1e-4


def test_full_size_instances_properties(ns, engine):
    """Config C5 at full size (1,048,576 instances): (i) the mix of the four rank shards (instance i on rank
    i % 4) summed equals the mix of the whole bank -- sharding does not change the result; (ii) with every gain
    but 64 picked instances' set to zero the million-voice render must equal the oracle mix of those 64."""
    from signals_b200.chain import ext
    n, frames = 1 << 20, RATE // 2
    whole_prm = cases.instance_params(5, n)
    compiled = engine.compile(cases.build_instances(ns, ext, whole_prm), 2, RATE)
    assert compiled.describe()['launches'][0]['channels_per_thread'] == 4
    whole = compiled.render_device(0, frames).cpu().numpy().astype(np.float64)
    compiled.close()
    parts = np.zeros_like(whole)
    for r in range(4):
        c = engine.compile(cases.build_instances(ns, ext, cases.instance_params(5, n, r, 4)), 2, RATE)
        parts += c.render_device(0, frames).cpu().numpy()
        c.close()
    assert np.isfinite(whole).all() and np.abs(whole).max() > 0.05
    assert max_abs_err(parts, whole) <= 1e-6
    rng = np.random.default_rng(9)
    pick = np.sort(rng.choice(n, 64, replace=False))
    prm = dict(whole_prm)
    prm['gain'] = np.zeros(n)
    prm['gain'][pick] = whole_prm['gain'][pick] * 256.0          # lift the 64 voices to a measurable level
    c = engine.compile(cases.build_instances(ns, ext, prm), 2, RATE)
    got = c.render_device(0, frames).cpu().numpy()
    c.close()
    sub = {k: (v[pick] if isinstance(v, np.ndarray) else v) for k, v in prm.items()}
    want = np_oracle.render_instances(sub, 0, frames, RATE)
    err = max_abs_err(got, want)
    print(f'C5 full size, 64 live voices of 1M: max-abs {err:.3e} (mix peak {np.abs(want).max():.3f})')
    assert err <= 1e-6


@pytest.mark.parametrize('name', ['sine_basic', 'square_edges', 'lowpass_c2_8v', 'highpass_order3', 'cascade8', 'mix', 'fanout'])
@pytest.mark.parametrize('frames', [1, 17, 1000])
def test_renders_stay_inside_their_block(name, frames, ns, engine):
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are caught with canaries: the block
    is a window (rows 3.., columns ..C) of a larger tensor whose every other element must stay untouched;
    also covers ragged row counts (1, 17) through every kernel family."""
    torch = _torch()
    case = cases.CASES_BY_NAME[name]
    c = case.channels
    big = torch.full((frames + 8, c + 5), -777.0, dtype=torch.float32, device='cuda')
    window = big[3:3 + frames, :c]
    compiled = engine.compile(case.build(ns), c, case.rate, frames)
    compiled.render_device(case.position, frames, window)
    compiled.close()
    host = big.cpu().numpy()
    inside = host[3:3 + frames, :c].copy()
    host[3:3 + frames, :c] = -777.0
    assert (host == -777.0).all(), 'a kernel wrote outside its (frames, channels) block'
    assert not (inside == -777.0).any()
    want = load_golden(name)
    if case.stride == 1 and frames <= len(want):
        err = max_abs_err(inside[np.isfinite(want[:frames]).all(1)], want[:frames][np.isfinite(want[:frames]).all(1)])
        assert err <= case.tol, err


def test_fused_reductions_stay_inside_their_block(ns, engine):
    from signals_b200.chain import ext
    torch = _torch()
    hertz, phase, amp = cases.bank_params(3, 300, 100)
    prm = cases.instance_params(5, 500)
    for graph, c in ((cases.build_bank(ns, ext, hertz, phase, amp, 3), 3), (cases.build_instances(ns, ext, prm), 2)):
        for frames in (1, 9, 700):
            big = torch.full((frames + 8, c + 5), -777.0, dtype=torch.float32, device='cuda')
            compiled = engine.compile(graph, c, RATE)
            compiled.render_device(5, frames, big[3:3 + frames, :c])
            compiled.close()
            host = big.cpu().numpy()
            assert not (host[3:3 + frames, :c] == -777.0).any()
            host[3:3 + frames, :c] = -777.0
            assert (host == -777.0).all()


def test_single_section_kernel_ragged_tiles_and_plain_stores(ns, engine):
    """k_chain_scan3 (64-channel tiles): channel counts that leave the second half of a tile partly or wholly
    dead, a Buffer source, and an output block whose row pitch defeats the TMA store (plain STG path)."""
    from signals_b200.chain import ext
    torch = _torch()
    rng = np.random.default_rng(46)
    for ch in (40, 70, 97):
        frames = 5000
        x = rng.uniform(-1, 1, (frames, ch)).astype(np.float32)
        cut = np.exp(rng.uniform(np.log(150.0), np.log(9000.0), (1, ch)))
        node = cases.lowpass(ns, ext.Buffer(x), [cut[0]], 'HighPass')
        want, _ = np_oracle.render_cascade(x.astype(np.float64), cut, RATE, btype='hp')
        compiled = engine.compile(node, ch, RATE)
        a = compiled.render_device(0, frames).cpu().numpy()                 # TMA stores when ch % 4 == 0
        big = torch.full((frames, ch + 1), 5.0, dtype=torch.float32, device='cuda')
        compiled.reset()
        compiled.render_device(0, frames, big[:, :ch])                       # pitch ch + 1: never TMA
        compiled.close()
        b = big.cpu().numpy()
        assert (b[:, ch] == 5.0).all()
        assert max_abs_err(a, want) <= 1e-4 and max_abs_err(b[:, :ch], want) <= 1e-4
        assert np.array_equal(a, b[:, :ch])


def test_instances_time_segments_match_oracle(ns, engine):
    """k_voices cuts a small bank along time as well (later segments warm their filters up from zero state over
    the decay horizon, oscillators need none): every segment must match the oracle, agree with the unsegmented
    render, and hand the true filter state to the next call."""
    from signals_b200.chain import ext
    prm = cases.instance_params(57, 600)
    prm['cutoff'] = np.clip(prm['cutoff'], 500.0, None)          # decay horizon ~1200 rows -> several segments fit
    frames = 60000
    compiled = engine.compile(cases.build_instances(ns, ext, prm), 2, RATE)
    first = compiled.render_device(0, frames).cpu().numpy()
    second = compiled.render_device(frames, 2000).cpu().numpy()
    launches_split = compiled.launch_count
    compiled.set_option('voices_segments', 1)
    compiled.reset()
    whole = compiled.render_device(0, frames).cpu().numpy()
    compiled.close()
    want = np_oracle.render_instances(prm, 0, frames + 2000, RATE)
    err = max_abs_err(np.concatenate([first, second]), want)
    print(f'instances time segments: max-abs {err:.3e} (mix peak {np.abs(want).max():.3f}), {launches_split} launches')
    assert err <= 1e-6
    assert max_abs_err(first, whole) <= 2e-7


def test_single_section_kernel_60s_low_cutoffs(ns, engine):
    """The default single-section kernel chains its sub-chunk carries in float32: 96 voices x 60 s with cutoffs
    down to 20 Hz (poles at radius 0.998) must stay inside the cascaded-IIR budget, and streaming the minute in
    ten calls must agree with the single request."""
    v, frames = 96, 60 * RATE
    hertz, phase, cutoff, g = cases.voice_params(88, v)
    cutoff[:32] = np.linspace(20.0, 100.0, 32)
    graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])
    compiled = engine.compile(graph, v, RATE)
    whole = compiled.render_device(0, frames).cpu().numpy()
    compiled.reset()
    parts = np.concatenate([compiled.render_device(k * frames // 10, frames // 10).cpu().numpy() for k in range(10)])
    compiled.close()
    pick = np.concatenate([np.arange(0, 32, 4), [40, 70, 95]])
    want = np_oracle.render_voice_chain(0, frames, RATE, hertz[pick], phase[pick], cutoff[pick], g[pick])
    err = max_abs_err(whole[:, pick], want)
    print(f'single-section kernel, 60 s, cutoffs from 20 Hz: max-abs {err:.3e}; streamed vs whole {max_abs_err(parts, whole):.3e}')
    assert err <= 1e-4
    assert max_abs_err(parts, whole) <= 2e-5


def test_wide_pointwise_nodes_vector_path(ns, engine):
    """Mix / RingMod / Amp / modulated Gain on 64-channel blocks take the 128-bit k_ewise_v4 path (the golden
    cases are 2-4 channels wide and stay on the scalar kernel); 63 channels fall back to it."""
    rng = np.random.default_rng(64)
    for c in (64, 63):
        hz_a, hz_b = rng.uniform(100, 2000, c), rng.uniform(50, 900, c)
        m = ns.Mix()
        m.left = cases.osc(ns, 'Sine', [hz_a], [rng.uniform(0, 1, c)])
        m.right = cases.osc(ns, 'Sawtooth', [hz_b])
        m.mix = cases.fixed(ns, [rng.uniform(0, 1, c)])
        r = ns.RingMod()
        r.left = m
        r.right = cases.osc(ns, 'Triangle', [rng.uniform(1, 20, c)], [rng.uniform(0, 1, c)])
        a = ns.Amp()
        a.left = r
        a.right = cases.fixed(ns, [rng.integers(1, 4, c).astype(float)])
        g = ns.Gain()
        g.left = a
        g.right = cases.osc(ns, 'Sine', [rng.uniform(0.5, 3.0, c)], [rng.uniform(0, 1, c)])     # modulated gain
        compiled = engine.compile(g, c, RATE)
        kinds = [l['kind'] for l in compiled.describe()['launches']]
        assert kinds == ['chain', 'chain', 'ewise', 'ewise']      # Mix and RingMod ride on oscillator chains as epilogues
        got = compiled.render_device(777, 3000).cpu().numpy()
        compiled.close()
        want = np_oracle.GraphOracle(RATE).render(g, 777, 3000, c)
        err = max_abs_err(got, want)
        print(f'Mix -> RingMod -> Amp -> Gain on {c} channels: max-abs {err:.3e}')
        # Amp raises to the power 1..3: |d(x^3)| = 3 x^2 |dx| triples the oscillators' 3.4e-7, hence 2e-6 and not 1e-6
        assert err <= 2e-6, c


def test_fused_pointwise_equals_materialised_path(ns, engine):
    """Mix / RingMod fused as the epilogue of a stateless oscillator chain (second operand: an oscillator kept in
    registers, or a block read once) against the same graphs run through k_ewise on materialised operands."""
    for name in ('mix', 'ringmod', 'unconnected', 'lfo_mix_amp', 'broadcast'):
        case = cases.CASES_BY_NAME[name]
        fused = render_case(engine, ns, case)
        _set_default('fuse_pointwise', 0)
        try:
            plain = render_case(engine, ns, case)
        finally:
            _set_default('fuse_pointwise', 1)
        ok = np.isfinite(plain)
        assert max_abs_err(fused[ok], plain[ok]) <= 5e-7, name
    # one side filtered (materialised), the other a bare oscillator
    m = ns.Mix()
    m.left = cases.lowpass(ns, cases.osc(ns, 'Sawtooth', [[220.0, 331.0]]), [[900.0, 2500.0]])
    m.right = cases.osc(ns, 'Sine', [[440.0]])                       # one channel, broadcast over two
    m.mix = cases.fixed(ns, [[0.3, 0.8]])
    compiled = engine.compile(m, 2, RATE)
    d = compiled.describe()['launches']
    assert [l['kind'] for l in d] == ['chain', 'chain'] and d[1]['epilogue'] == 'mix' and d[1]['other'] == 'block'
    got = compiled.render_device(0, 4800).cpu().numpy()
    compiled.close()
    assert max_abs_err(got, np_oracle.GraphOracle(RATE).render(m, 0, 4800, 2)) <= 1e-4


def test_modulated_cutoff_streams_with_carried_state(ns, engine):
    """Contiguous requests through a filter whose cutoff is driven by an LFO: the sections are re-designed on the
    device at the first frame of every request (k_design) and the integrator states carry over.  The reference has
    no streaming equivalent (it restarts every block from zero state); the expectation is the same
    state-variable section in float64 with the reference's per-request cutoff sampling (fx.py:124-129)."""
    src = cases.osc(ns, 'Sawtooth', [[220.0, 331.0]])
    wah = cases._wah(ns, [[400.0, 900.0]], [[3000.0, 5200.0]], [[7.0, 11.0]], [[0.1, 0.4]])
    node = cases._with_cutoff(ns, src, wah)
    block, nblocks = 512, 12
    compiled = engine.compile(node, 2, RATE)
    got = np.concatenate([compiled.render_device(b * block, block).cpu().numpy() for b in range(nblocks)])
    compiled.close()
    orc = np_oracle.GraphOracle(RATE)
    x = orc.render(src, 0, block * nblocks, 2)
    want = np.zeros_like(x)
    r2 = 2.0 * np.sin(np.pi / 4.0)
    cutoffs = []
    for c in range(2):
        s1 = s2 = 0.0
        for b in range(nblocks):
            fc = float(np.broadcast_to(orc.render(wah, b * block, 1, 2), (1, 2))[0, c])
            if c == 0:
                cutoffs.append(fc)
            g = np.tan(np.pi * fc / RATE)
            d = 1.0 / (1.0 + r2 * g + g * g)
            for n in range(b * block, (b + 1) * block):
                hp = (x[n, c] - (r2 + g) * s1 - s2) * d
                bp = g * hp + s1
                s1 = g * hp + bp
                lp = g * bp + s2
                s2 = g * bp + lp
                want[n, c] = lp
    assert max(cutoffs) - min(cutoffs) > 1000.0            # the sweep really moves the filter between requests
    assert max_abs_err(got, want) <= 1e-4


@pytest.mark.parametrize('wave,nsec,btype,ch', [('Sine', 1, 'lp', 128), ('Sine', 1, 'lp', 130), ('Square', 2, 'hp', 66),
                                                ('Sawtooth', 8, 'lp', 64), ('Triangle', 4, 'lp', 5), ('Sine', 2, 'lp', 128),
                                                ('Sawtooth', 8, 'hp', 64), ('Square', 3, 'mix', 70), ('Sine', 6, 'mix', 64),
                                                ('Triangle', 4, 'svf', 64)])
def test_osc_reg_pieces_and_streaming(wave, nsec, btype, ch, ns, engine):
    """k_osc_delta / k_osc_reg (oscillator evaluated in the thread, all sections in registers; 'mix' alternates high- and
    low-pass sections, 'svf' keeps k_osc_reg's state-variable sections): time pieces with decay warm-up against
    the float64 reference render, agreement with the uncut render, state carried into a ragged second request, and
    agreement of a chunked render with the single request."""
    rng = np.random.default_rng(46)
    frames, tail = 60000, 3001
    cut = np.exp(rng.uniform(np.log(700.0), np.log(8000.0), (nsec, ch)))
    node = cases.osc(ns, wave, [rng.uniform(55.0, 1760.0, ch)], [rng.uniform(0, 1, ch)])
    kind = lambda s: 'HighPass' if btype == 'hp' or (btype == 'mix' and s % 2 == 0) else 'LowPass'     # noqa: E731
    for s in range(nsec):
        node = cases.lowpass(ns, node, [cut[s]], kind(s))
    compiled = engine.compile(node, ch, RATE)
    compiled.set_option('osc_reg', 1)
    if btype == 'svf':
        compiled.set_option('osc_delta', 0)
    warm = compiled.describe()['launches'][0]['warm_rows']
    assert 0 < warm < frames // 8, warm                       # so that the launch really is cut along time
    first = compiled.render_device(0, frames).cpu().numpy()
    second = compiled.render_device(frames, tail).cpu().numpy()
    compiled.set_option('pipe_segments', 1)
    compiled.reset()
    whole = compiled.render_device(0, frames).cpu().numpy()
    compiled.reset()
    cuts = [0, 1, 9, 1000, 1016, 4803, 20000]
    parts = np.concatenate([compiled.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
    compiled.close()
    pick = rng.choice(ch, min(ch, 12), replace=False)
    sub = cases.osc(ns, wave, [np.asarray(node_hertz(node))[pick]], [np.asarray(node_phase(node))[pick]])
    for s in range(nsec):
        sub = cases.lowpass(ns, sub, [cut[s][pick]], kind(s))
    want = np_oracle.GraphOracle(RATE).render(sub, 0, frames + tail, len(pick))
    err = max_abs_err(np.concatenate([first, second])[:, pick], want)
    print(f'osc_reg {wave} x {nsec} {btype} sections, {ch} ch: warm_rows {warm}, max-abs {err:.3e}')
    assert err <= 1e-4
    assert max_abs_err(first, whole) <= 1e-6
    # (not bit-equal: inside a block the phase word advances by the rounded increment, and the blocks of a chunked
    # render start at other rows)
    assert max_abs_err(parts, whole[:20000]) <= 2e-6


def _osc_of(node):
    while 'input' in node.inputs_by_port:
        node = node.inputs_by_port['input']
    return node


def node_hertz(node):
    return _osc_of(node).inputs_by_port['hertz'].get_state().value[0]


def node_phase(node):
    return _osc_of(node).inputs_by_port['phase'].get_state().value[0]
