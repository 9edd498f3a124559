"""GPU: the round-2 paths, all through the C ABI -- the realtime block path (sigb_render_block: captured CUDA graph,
pinned header / staging), taps read back from the same launch, block-rate parameters sampled once per request by
sigb_render_host, modulated cutoffs on the time-parallel kernels (k_design writes their tables), the device-side
"cutoff outside (0, Nyquist)" error, seek warm-up in slab-sized pieces, and the equal-piece decomposition of k_voices."""
import numpy as np
import pytest

from conftest import load_golden, max_abs_err
from oracle import cases, np_oracle

pytestmark = pytest.mark.gpu

RATE = 48000


def _blocks(compiled, frames, block, how):
    out = np.empty((frames, compiled.channels), dtype=np.float32)
    for r in range(0, frames, block):
        n = min(block, frames - r)
        if how == 'block':
            compiled.render_block(r, n, out[r:r + n])
        else:
            compiled.render_host(r, n, out[r:r + n])
    return out


@pytest.mark.parametrize('name', ['lowpass_c2_8v', 'cascade8', 'mix', 'lfo_chain', 'lfo_cutoff_cascade', 'fanout', 'highpass_order3'])
@pytest.mark.parametrize('graph', [1, 0])
def test_render_block_equals_render_host(name, graph, ns, engine):
    """The audio-callback path: 384-frame blocks through sigb_render_block (one captured CUDA graph launch per
    block, or the same launches issued directly) carry the filter state and re-sample the block-rate parameters at every
    block exactly as sigb_render_host does for the same sequence of requests."""
    case = cases.CASES_BY_NAME[name]
    frames, block = 384 * 9 + 100, 384
    a = engine.compile(case.build(ns), case.channels, RATE)
    a.set_option('force_seq', 1)          # the block path runs one sequential launch per chain: compare like with like
    want = _blocks(a, frames, block, 'host')
    a.close()
    b = engine.compile(case.build(ns), case.channels, RATE)
    b.set_option('rt_graph', graph)
    got = _blocks(b, frames, block, 'block')
    g = b.graph_launches
    b.close()
    assert np.array_equal(got, want), f'{name}: max-abs {max_abs_err(got, want):.3e}'
    # first block = seek (served by render_host), then one graph launch per full block and per ragged last block
    assert g == (frames // block if graph else 0)


def test_render_block_seek_and_recapture(ns, engine):
    """A seek falls back to the host path (zero state + context warm-up) and the blocks after it replay the graph; a
    device render in between moves the live copy of the double-buffered state and the graph is re-captured."""
    case = cases.CASES_BY_NAME['lowpass_c2_8v']
    c = engine.compile(case.build(ns), case.channels, RATE)
    ref = engine.compile(case.build(ns), case.channels, RATE)
    out = np.empty((512, case.channels), np.float32)
    want = np.empty_like(out)
    for pos in (0, 512, 1024, 40000, 40512):
        c.render_block(pos, 512, out)
        ref.render_host(pos, 512, want)
        assert max_abs_err(out, want) <= 2e-6, pos
    dev = c.render_device(41024, 4096).cpu().numpy()            # time-parallel kernel: flips the state copy
    wdev = ref.render_device(41024, 4096).cpu().numpy()
    assert max_abs_err(dev, wdev) <= 2e-6
    for pos in (45120, 45632):
        c.render_block(pos, 512, out)
        ref.render_host(pos, 512, want)
        assert max_abs_err(out, want) <= 2e-6, pos
    c.close()
    ref.close()


def test_sink_device_uses_the_graph_path_and_the_dirty_flag(ns):
    """SinkDevice.render_block: the plan is kept while the graph epoch stands still (no walk), every contiguous
    callback is one CUDA graph launch, an edited parameter recompiles, and an in-place write into a Fixed's array is
    noticed too."""
    from signals_b200 import engine as engine_mod
    from signals_b200.chain import dev
    info = dev.DeviceInfo(name='t', index=0, hostapi=0, max_input_channels=0, max_output_channels=2, default_low_input_latency=0.0,
                          default_low_output_latency=0.0, default_high_input_latency=0.0, default_high_output_latency=0.0,
                          default_samplerate=float(RATE))
    eng = engine_mod.default_engine()
    eng.clear()
    sink = dev.SinkDevice(info)
    hz = cases.fixed(ns, [[500.0]])
    osc = ns.Sine()
    osc.hertz = hz
    sink.input = cases.gain(ns, osc, [[0.2]])
    out = np.zeros((512, 1), np.float32)
    blocks = []
    for _ in range(6):
        sink.render_block(out, 512, RATE)
        blocks.append(out.copy())
    compiled = eng.plan_for(sink._ports['input'].sig, 1, RATE, 512)
    assert compiled.graph_launches == 5
    want = np_oracle.example_sine_block(0, 6 * 512, RATE)
    assert max_abs_err(np.concatenate(blocks), want) <= 1e-6
    hz.get_state().value = np.array([[250.0]])                       # setter -> epoch -> recompile
    sink.render_block(out, 512, RATE)
    assert eng.plan_for(sink._ports['input'].sig, 1, RATE, 512) is not compiled
    assert max_abs_err(out, np_oracle.example_sine_block(6 * 512, 512, RATE, frequency=250.0)) <= 1e-6
    hz.get_state().value[0, 0] = 125.0                               # in-place write: caught by the snapshot compare
    sink.render_block(out, 512, RATE)
    assert max_abs_err(out, np_oracle.example_sine_block(7 * 512, 512, RATE, frequency=125.0)) <= 1e-6
    eng.clear()


def test_interior_taps_are_served_by_the_same_launch(ns):
    """Wave taps inside the graph: their blocks come from the buffers the one render left in HBM (sigb_plan_read_tap),
    not from a second render; a tap at the root gets the rendered block itself."""
    from signals_b200 import engine as engine_mod
    from signals_b200.chain import vis
    eng = engine_mod.Engine()
    src = cases.osc(ns, 'Sawtooth', [[220.0, 330.0]])
    t1 = vis.Wave()
    t1.input = src
    lp = cases.lowpass(ns, t1, [[900.0, 1500.0]])
    t2 = vis.Wave()
    t2.input = cases.gain(ns, lp, [[0.5, 0.25]])
    from signals_b200.chain import BlockLoc, Shape
    compiled = eng.plan_for(t2, 2, RATE, 1000)
    kinds = [(l['kind'], l.get('sections')) for l in compiled.describe()['launches']]
    assert kinds == [('chain', 0), ('chain', 1)]                       # the interior tap keeps the oscillator block aside
    launches = []
    for pos in (0, 1000):
        loc = BlockLoc(position=pos, rate=RATE, shape=Shape(frames=1000, channels=2))
        block = eng.render(t2, loc)
        n0 = compiled.launch_count
        assert eng.serve_taps(t2, loc, rendered=block) == 2
        launches.append(compiled.launch_count - n0)
    assert launches == [0, 0]                                          # no kernel ran for the taps
    orc = np_oracle.GraphOracle(RATE)
    got_src = np.concatenate([t1.q.get(), t1.q.get()])
    got_out = np.concatenate([t2.q.get(), t2.q.get()])
    assert max_abs_err(got_src, orc.render(src, 0, 2000, 2)) <= 1e-6
    plain = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sawtooth', [[220.0, 330.0]]), [[900.0, 1500.0]]), [[0.5, 0.25]])
    assert max_abs_err(got_out, orc.render(plain, 0, 2000, 2)) <= 1e-4           # the same graph without the taps
    eng.clear()


@pytest.mark.parametrize('name', ['lfo_hertz', 'lfo_gain', 'lfo_cutoff', 'lfo_mix_amp'])
def test_render_host_samples_block_rate_parameters_once_per_request(name, ns, engine):
    """One request = one sampling of the modulated parameters, at its first frame (forward_at_block_rate,
    chain/__init__.py:305-306), however sigb_render_host cuts the request into staging slabs."""
    case = cases.CASES_BY_NAME[name]
    a = engine.compile(case.build(ns), case.channels, RATE)
    want = a.render_device(case.position, case.frames).cpu().numpy()
    a.close()
    b = engine.compile(case.build(ns), case.channels, RATE)
    b.set_option('host_slab_bytes', 4 * case.channels * 500)        # ~500-row host slabs
    got = b.render_host(case.position, case.frames)
    b.close()
    assert max_abs_err(got, want) <= 1e-6, name
    assert max_abs_err(got[::case.stride], load_golden(name)) <= case.tol


@pytest.mark.parametrize('name', ['lfo_cutoff', 'lfo_cutoff_hp3', 'lfo_cutoff_cascade'])
def test_modulated_cutoffs_run_on_the_time_parallel_kernels(name, ns, engine):
    """k_design writes the scan tables and the decay horizon of the request's design, so a chain with an LFO on its
    cutoff takes the same time-parallel kernels as a constant one: same launches as the unmodulated chain, same result as
    the sequential kernel, and the reference's golden."""
    case = cases.CASES_BY_NAME[name]
    frames = max(case.frames, 48000)
    seq = engine.compile(case.build(ns), case.channels, RATE)
    seq.set_option('force_seq', 1)
    want = seq.render_device(case.position, frames).cpu().numpy()
    seq.close()
    par = engine.compile(case.build(ns), case.channels, RATE)
    got = par.render_device(case.position, frames).cpu().numpy()
    n_par = par.launch_count
    par.close()
    assert max_abs_err(got, want) <= 2e-5, name
    assert max_abs_err(got[:case.frames:case.stride], load_golden(name)) <= case.tol
    # param eval + one design per modulated filter (x2: seek samples the context position too) + the chain kernels
    assert n_par >= 3


def test_modulated_cutoff_time_pieces_match_the_float64_section(ns, engine):
    """A long request with a modulated cutoff on 256 channels: the scan kernel cuts tiles along time with the decay
    horizon k_design reported; compare with the float64 state-variable section at the sampled cutoff."""
    ch, frames = 256, 96000
    rng = np.random.default_rng(11)
    hz = rng.uniform(60.0, 2000.0, ch)
    src = cases.osc(ns, 'Sine', [hz])
    lo, hi = rng.uniform(300.0, 600.0, ch), rng.uniform(2000.0, 6000.0, ch)
    wah = cases._wah(ns, [lo], [hi], [rng.uniform(0.5, 3.0, ch)], [rng.uniform(0.0, 1.0, ch)])
    node = cases._with_cutoff(ns, src, wah)
    c = engine.compile(node, ch, RATE)
    pos = 4800
    got = c.render_device(pos, frames).cpu().numpy()
    c.close()
    want = np_oracle.GraphOracle(RATE).render(node, pos, frames, ch)
    # the oracle restarts 100 frames before the request from zero state, as the plan does on a seek
    assert max_abs_err(got, want) <= 1e-4


def test_modulated_cutoff_outside_nyquist_raises_like_scipy(ns, engine):
    """A cutoff driven to <= 0 Hz: scipy.signal.butter raises ValueError in the reference (fx.py:102); the device-side
    design flags it and the call that finds the flag raises the same error class."""
    from signals_b200.chain import FilterDesignError
    src = cases.osc(ns, 'Sine', [[440.0]])
    lfo = cases.osc(ns, 'Sine', [[0.25]], [[0.75]])                 # -1 at position 0
    node = cases._with_cutoff(ns, src, cases.gain(ns, lfo, [[1000.0]]))
    c = engine.compile(node, 1, RATE)
    out = np.empty((256, 1), np.float32)
    with pytest.raises(FilterDesignError):
        c.render_host(0, 256, out)
    with pytest.raises(ValueError):
        np_oracle.GraphOracle(RATE).render(node, 0, 256, 1)
    c.close()


def test_seek_warmup_runs_in_slab_sized_pieces(ns, engine):
    """A device slab shorter than the seek's context (64 rows against 800 frames of context for 8 chained filters): the
    warm-up runs slab by slab inside buffers of one slab (canary rows behind the output stay untouched)."""
    import torch
    case = cases.CASES_BY_NAME['cascade8']
    src = cases.gain(ns, case.build(ns), [[1.0] * case.channels])
    mix = ns.Mix()                      # forces materialised intermediates (plan buffers)
    mix.left = src
    mix.right = cases.osc(ns, 'Sine', [[100.0] * case.channels])
    mix.mix = cases.fixed(ns, [[0.5] * case.channels])
    lp = cases.lowpass(ns, mix, [[2000.0] * case.channels])
    a = engine.compile(lp, case.channels, RATE)
    want = a.render_device(5000, 700).cpu().numpy()
    a.close()
    b = engine.compile(lp, case.channels, RATE)
    b.set_option('slab_frames', 64)
    out = torch.full((700 + 64, case.channels), 7.0, device='cuda')
    b.render_device(5000, 700, out)
    got = out.cpu().numpy()
    b.close()
    assert np.all(got[700:] == 7.0)
    assert max_abs_err(got[:700], want) <= 2e-6


@pytest.mark.parametrize('n,frames,pieces', [(9000, 24000, 0), (9000, 24000, 7), (2500, 30000, 0), (40, 5000, 0)])
def test_voice_bank_pieces_match_oracle(n, frames, pieces, ns, engine):
    """k_voices cuts the (voice group, row block) space into equal pieces per CTA slot: any bank size, pieces that
    start inside a group (decay warm-up), pieces that span several groups, and the state handed to the next call."""
    from signals_b200.chain import ext
    prm = cases.instance_params(77, n)
    prm['cutoff'] = np.clip(prm['cutoff'], 600.0, None)              # decay horizon ~1000 rows: cuts inside groups fit
    compiled = engine.compile(cases.build_instances(ns, ext, prm), 2, RATE)
    if pieces:
        compiled.set_option('voices_pieces', pieces)
    first = compiled.render_device(0, frames).cpu().numpy()
    second = compiled.render_device(frames, 1000).cpu().numpy()
    compiled.set_option('voices_segments', 1)                        # one piece per group: no cuts along time
    compiled.reset()
    whole = compiled.render_device(0, frames).cpu().numpy()
    compiled.close()
    want = np_oracle.render_instances(prm, 0, frames + 1000, RATE)
    err = max_abs_err(np.concatenate([first, second]), want)
    assert err <= 1e-6, err
    assert max_abs_err(first, whole) <= 2e-7


def test_blockwise_reference_option_is_what_separates_the_two_stream_semantics(ns, engine):
    """Golden lowpass_blockwise_stream = the reference driven block by block (every request restarts the filter 100
    frames early).  blockwise_reference=1 reproduces it (checked by the golden test); the default carries the true state,
    which is the single-request render (the oracle) and differs from the blockwise reference by a few 1e-4."""
    case = cases.CASES_BY_NAME['lowpass_blockwise_stream']
    c = engine.compile(case.build(ns), case.channels, RATE)
    c.render_device(0, case.position)                                   # stream from 0 up to the first block
    carried = np.concatenate([c.render_device(case.position + r, case.block).cpu().numpy() for r in range(0, case.frames, case.block)])
    c.close()
    hertz, phase, cutoff, g = cases.voice_params(7, 4)
    single = np_oracle.render_voice_chain(0, case.position + case.frames, RATE, hertz, phase, cutoff, g)[case.position:]
    assert max_abs_err(carried, single) <= 1e-5
    d = max_abs_err(carried, load_golden(case.name))
    print(f'carried state vs the reference\'s blockwise render: {d:.3e}')
    assert 1e-5 < d < 5e-3


# ------------------------------------------------------------------------------------------------
# BASELINE full sizes against float64 oracle renders of the WHOLE blocks (fixtures: oracle/make_fullsize.py).  A cell of
# the grids sums 1,000 rows x 64 channels: it moves by tens to hundreds when a single 64-channel x 16-row tile of the
# block is wrong and by < 0.1 under the per-sample budgets, so every tile of the block is covered, not a sample of voices.
# ------------------------------------------------------------------------------------------------
def _load(name, key):
    import os
    from conftest import GOLDEN_DIR
    with np.load(os.path.join(GOLDEN_DIR, name)) as z:
        return z[key]


def _grid(block, rows=1000, ch=64):
    f, c = block.shape
    return block.double().reshape(f // rows, rows, c // ch, ch).sum(dim=(1, 3)).cpu().numpy()


def test_full_size_c2_every_tile_against_the_oracle(ns, engine):
    v, frames = 4096, 10 * RATE
    hertz, phase, cutoff, g = cases.voice_params(2, v)
    graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sine', [hertz], [phase]), [cutoff]), [g])
    compiled = engine.compile(graph, v, RATE)
    got = _grid(compiled.render_device(0, frames))
    compiled.close()
    want = _load('full_c2_grid.npz', 'grid')
    dev = np.abs(got - want)
    print(f'C2 full block, {want.size} cells of 1000 x 64: max cell deviation {dev.max():.3e} (cells up to {np.abs(want).max():.1f})')
    assert dev.max() <= 0.05


def test_full_size_c4_every_tile_against_the_oracle(ns, engine):
    import torch
    from oracle.make_fullsize import c4_noise
    from signals_b200.chain import ext
    ch, frames = 16384, RATE
    rng = np.random.default_rng(4)
    cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, ch)))
    node = ext.Buffer(torch.from_numpy(c4_noise(frames, ch)).cuda())
    for s in range(8):
        node = cases.lowpass(ns, node, [cut[s]])
    compiled = engine.compile(node, ch, RATE, frames)
    got = _grid(compiled.render_device(0, frames))
    compiled.close()
    want = _load('full_c4_grid.npz', 'grid')
    dev = np.abs(got - want)
    print(f'C4 full width, first second, {want.size} cells of 1000 x 64: max cell deviation {dev.max():.3e} (cells up to {np.abs(want).max():.1f})')
    assert dev.max() <= 0.05


def test_full_size_c5_mix_against_the_oracle(ns, engine):
    from signals_b200.chain import ext
    n, frames = 1 << 20, 4800
    compiled = engine.compile(cases.build_instances(ns, ext, cases.instance_params(5, n)), 2, RATE)
    got = compiled.render_device(0, frames).cpu().numpy()
    compiled.close()
    want = _load('full_c5_mix.npz', 'mix')
    err = max_abs_err(got, want)
    print(f'C5, all 1,048,576 instances, first {frames} frames of the mix: max-abs {err:.3e} (peak {np.abs(want).max():.3f})')
    assert err <= 1e-6


# ------------------------------------------------------------------------------------------------
# edges: empty requests, very wide blocks, ragged tiles on every path
# ------------------------------------------------------------------------------------------------
def test_empty_requests_are_no_ops_on_every_entry_point(ns, engine):
    """A request of zero frames (BlockLoc with shape (0, C)) renders nothing, launches nothing and does not disturb the
    stream position: the next block continues the carried state."""
    import torch
    case = cases.CASES_BY_NAME['lowpass_c2_8v']
    c = engine.compile(case.build(ns), case.channels, RATE)
    ref = engine.compile(case.build(ns), case.channels, RATE)
    first = c.render_device(0, 300).cpu().numpy()
    n0 = c.launch_count
    c.render_device(300, 0, torch.empty((0, case.channels), device='cuda'))
    c.render_host(300, 0, np.empty((0, case.channels), np.float32))
    c.render_block(300, 0, np.empty((0, case.channels), np.float32))
    assert c.launch_count == n0
    second = c.render_device(300, 300).cpu().numpy()
    want = ref.render_device(0, 600).cpu().numpy()
    assert max_abs_err(np.concatenate([first, second]), want) <= 1e-6
    c.close()
    ref.close()


def test_very_wide_blocks(ns, engine):
    """A million-channel oscillator block and a 100,003-channel filtered block (ragged last tiles of every kernel): spot
    checks against the oracle plus finiteness of the whole block."""
    import torch
    rng = np.random.default_rng(123)
    c = 1_000_003
    hz, ph = rng.uniform(20.0, 12000.0, c), rng.uniform(0.0, 1.0, c)
    comp = engine.compile(cases.osc(ns, 'Sine', [hz], [ph]), c, RATE)
    out = comp.render_device(1000, 48)
    comp.close()
    assert bool(torch.isfinite(out).all())
    pick = np.concatenate([[0, 1, c - 2, c - 1], rng.choice(c, 60, replace=False)])
    got = out[:, torch.from_numpy(pick).cuda()].cpu().numpy()
    want = np_oracle.sine(np_oracle.osc_cycles(1000, 48, RATE, hz[pick].reshape(1, -1), ph[pick].reshape(1, -1)))
    assert max_abs_err(got, want) <= 1e-6
    del out
    c = 100_003
    hz, ph, cut, g = cases.voice_params(9, c)
    graph = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, 'Sawtooth', [hz], [ph]), [cut]), [g])
    comp = engine.compile(graph, c, RATE)
    out = comp.render_device(0, 1500)                 # 10 scan steps + a 60-row tail
    comp.close()
    assert bool(torch.isfinite(out).all())
    pick = np.concatenate([[0, 63, 64, c - 4, c - 1], rng.choice(c, 40, replace=False)])
    got = out[:, torch.from_numpy(pick).cuda()].cpu().numpy()
    want = np_oracle.render_voice_chain(0, 1500, RATE, hz[pick], ph[pick], cut[pick], g[pick], wave='Sawtooth')
    assert max_abs_err(got, want) <= 1e-4


@pytest.mark.parametrize('cls,btype,order,nodes', [('LowPass', 'lp', 3, 3), ('HighPass', 'hp', 5, 1), ('LowPass', 'lp', 5, 2)])
def test_odd_order_cascades_on_the_register_kernel(cls, btype, order, nodes, ns, engine):
    """Odd Butterworth orders end in a first-order section.  k_cascade_reg / k_osc_reg run it on the second-order
    section's six instructions with coefficients chosen at load time (c = 1, g d = G, 2g = 0; low-pass g = 1, high-pass
    d = 1 - G), so such cascades no longer fall back to the section-pipelined kernel: same result as that kernel and as the
    float64 cascade, state carried into a second request, time pieces included."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(71)
    ch, frames = 192, 40000
    x = rng.uniform(-1, 1, (frames + 3000, ch)).astype(np.float32)
    cut = np.exp(rng.uniform(np.log(700.0), np.log(8000.0), (nodes, ch)))

    def graph():
        node = ext.Buffer(x)
        for s in range(nodes):
            node = cases.lowpass(ns, node, [cut[s]], cls, order)
        return node

    got = {}
    for kernel in ('reg', 'pipe'):
        c = engine.compile(graph(), ch, RATE)
        c.set_option('cascade_reg', -1 if kernel == 'reg' else 0)
        (launch,) = c.describe()['launches']
        assert launch['sections'] == nodes * ((order + 1) // 2)
        first = c.render_device(0, frames).cpu().numpy()
        second = c.render_device(frames, 3000).cpu().numpy()
        c.close()
        got[kernel] = np.concatenate([first, second])
    pick = rng.choice(ch, 16, replace=False)
    want, _ = np_oracle.render_cascade(x[:, pick].astype(np.float64), cut[:, pick], RATE, btype, order)
    err = max_abs_err(got['reg'][:, pick], want)
    print(f'{nodes} x {cls} order {order} on the register kernel: max-abs {err:.3e}; vs the pipelined kernel {max_abs_err(got["reg"], got["pipe"]):.3e}')
    assert err <= 1e-4
    assert max_abs_err(got['reg'], got['pipe']) <= 2e-6


def test_odd_order_oscillator_chain_on_the_register_kernel(ns, engine):
    """Oscillator -> three order-3 high-pass nodes (6 sections, every second one first-order) through k_osc_reg."""
    ch, frames = 130, 30000
    hz, ph, cut, g = cases.voice_params(73, ch)
    node = cases.osc(ns, 'Sawtooth', [hz], [ph])
    for k in range(3):
        node = cases.lowpass(ns, node, [np.clip(cut * (1.0 + 0.3 * k), 300.0, 9000.0)], 'HighPass', 3)
    c = engine.compile(node, ch, RATE)
    got = c.render_device(0, frames).cpu().numpy()
    c.close()
    pick = np.random.default_rng(1).choice(ch, 12, replace=False)
    orc = np_oracle.GraphOracle(RATE)
    sub = cases.osc(ns, 'Sawtooth', [hz[pick]], [ph[pick]])
    for k in range(3):
        sub = cases.lowpass(ns, sub, [np.clip(cut[pick] * (1.0 + 0.3 * k), 300.0, 9000.0)], 'HighPass', 3)
    want = orc.render(sub, 0, frames, len(pick))
    err = max_abs_err(got[:, pick], want)
    print(f'osc -> 3 x HighPass order 3: max-abs {err:.3e}')
    assert err <= 1e-4


@pytest.mark.parametrize('kinds', ['HHHHLLLL', 'LHLHHLLH', 'LLH', 'HLL', 'HLLLLLL', 'LHHHH'])
def test_mixed_cascades_on_the_register_kernel(kinds, ns, engine):
    """Cascades that MIX low- and high-pass sections (a band-pass built from chained HighPass / LowPass nodes) run
    register-resident in k_cascade_delta -- both kinds on the shared delta-form states, a select per section -- instead of
    the section-pipelined kernel: against scipy's float64 cascade, against k_cascade_pipe on the same plan, cut into time
    pieces, streamed in ragged requests, and handing its state to a second call."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(81)
    ch, frames = 128, 48000
    x = rng.uniform(-1, 1, (frames + 3000, ch)).astype(np.float32)
    cut = np.exp(rng.uniform(np.log(500.0), np.log(8000.0), (len(kinds), ch)))

    def graph():
        node = ext.Buffer(x)
        for s, k in enumerate(kinds):
            node = cases.lowpass(ns, node, [cut[s]], 'HighPass' if k == 'H' else 'LowPass')
        return node

    got = {}
    for kernel in ('reg', 'pipe'):
        c = engine.compile(graph(), ch, RATE)
        c.set_option('cascade_reg', -1 if kernel == 'reg' else 0)
        first = c.render_device(0, frames).cpu().numpy()
        second = c.render_device(frames, 3000).cpu().numpy()
        got[kernel] = np.concatenate([first, second])
        if kernel == 'reg':
            c.set_option('pipe_segments', 1)                  # never cut along time
            c.reset()
            whole = c.render_device(0, frames).cpu().numpy()
            c.reset()
            cuts = [0, 1, 17, 1000, 1016, 4803, 20000, frames]
            parts = np.concatenate([c.render_device(a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
            assert max_abs_err(first, whole) <= 1e-6
            assert max_abs_err(parts, whole) <= 5e-7          # the low-pass sections' one-row memory: rounding noise only
        c.close()
    want = x.astype(np.float64)
    pick = rng.choice(ch, 12, replace=False)
    want = want[:, pick]
    for s, k in enumerate(kinds):
        want, _ = np_oracle.render_cascade(want, cut[s:s + 1, pick], RATE, btype='hp' if k == 'H' else 'lp')
    err = max_abs_err(got['reg'][:, pick], want)
    print(f'mixed cascade {kinds} on the register kernel: max-abs {err:.3e}; vs the pipelined kernel {max_abs_err(got["reg"], got["pipe"]):.3e}')
    assert err <= 1e-4
    assert max_abs_err(got['reg'], got['pipe']) <= 2e-5


def test_streams_pass_between_the_cascade_kernels(ns, engine):
    """One stream, four requests, a different cascade kernel for each (delta form, state-variable sections in 8- and 4-row
    blocks, the section-pipelined kernel): the delta form's states are the state-variable section's own, re-scaled, so the
    hand-over is a pointwise conversion and the stream continues as if one kernel had rendered it."""
    from signals_b200.chain import ext
    rng = np.random.default_rng(82)
    ch, nsec, seg = 128, 8, 12000
    x = rng.uniform(-1, 1, (4 * seg, ch)).astype(np.float32)
    for cls, btype in (('LowPass', 'lp'), ('HighPass', 'hp')):
        cut = np.exp(rng.uniform(np.log(150.0), np.log(6000.0), (nsec, ch)))
        node = ext.Buffer(x)
        for s in range(nsec):
            node = cases.lowpass(ns, node, [cut[s]], cls)
        c = engine.compile(node, ch, RATE)
        out = []
        for k, (reg, variant) in enumerate([(-1, 0), (-1, 4), (0, 0), (-1, 0)]):
            c.set_option('cascade_reg', reg)
            c.set_option('reg_variant', variant)
            out.append(c.render_device(k * seg, seg).cpu().numpy())
        c.close()
        got = np.concatenate(out)
        want, _ = np_oracle.render_cascade(x.astype(np.float64), cut, RATE, btype=btype)
        err = max_abs_err(got, want)
        print(f'{cls} stream through delta / state-variable / pipelined / delta kernels: max-abs {err:.3e}')
        assert err <= 2e-6


@pytest.mark.parametrize('kinds', ['LL', 'HH', 'HL'])
def test_two_section_chains_on_a_block_leave_the_scan_kernel(kinds, ns, engine):
    """Two second-order sections on a materialised block (two chained filter nodes, or one order-4 node): when the time
    pieces their decay horizon allows fill the machine the chain runs in k_cascade_delta instead of k_chain_scan2 --
    same result (against the float64 cascade and against the scan kernel on the same plan), state carried on."""
    from signals_b200.chain import ext
    torch = pytest.importorskip('torch')
    rng = np.random.default_rng(83)
    ch, frames, tail = 4096, 96000, 1000
    g = torch.Generator(device='cuda')
    g.manual_seed(83)
    xd = torch.rand((frames + tail, ch), generator=g, device='cuda', dtype=torch.float32) * 2 - 1
    cut = np.exp(rng.uniform(np.log(800.0), np.log(8000.0), (2, ch)))
    got = {}
    for kernel in ('reg', 'scan'):
        node = ext.Buffer(xd)
        for s, k in enumerate(kinds):
            node = cases.lowpass(ns, node, [cut[s]], 'HighPass' if k == 'H' else 'LowPass')
        c = engine.compile(node, ch, RATE)
        c.set_option('cascade_reg', -1 if kernel == 'reg' else 0)
        first = c.render_device(0, frames)
        second = c.render_device(frames, tail)
        got[kernel] = torch.cat([first, second])
        c.close()
    pick = np.sort(rng.choice(ch, 10, replace=False))
    want = xd[:, torch.from_numpy(pick).cuda()].cpu().numpy().astype(np.float64)
    for s, k in enumerate(kinds):
        want, _ = np_oracle.render_cascade(want, cut[s:s + 1, pick], RATE, btype='hp' if k == 'H' else 'lp')
    err = max_abs_err(got['reg'][:, torch.from_numpy(pick).cuda()].cpu().numpy(), want)
    diff = float((got['reg'] - got['scan']).abs().max())
    print(f'two sections {kinds} on a block: register kernel max-abs {err:.3e}; vs the scan kernel {diff:.3e}')
    assert err <= 1e-4
    assert diff <= 2e-6


def test_large_modulated_two_section_request_runs_register_resident(ns, engine):
    """Two filters with LFO-driven cutoffs behind an oscillator in a LARGE request (>= 2^28 samples): the plan waits for the
    decay horizon k_design reports and takes k_osc_delta, like an unmodulated chain -- same result as the scan kernel
    (cascade_pipe = 0) and as the float64 oracle at the sampled cutoffs."""
    torch = pytest.importorskip('torch')
    ch, frames, pos = 4096, 144000, 0       # (position 0: a seek into CHAINED filters follows DESIGN 1, not the reference's nested restarts)
    rng = np.random.default_rng(84)
    hz = rng.uniform(60.0, 2000.0, ch)

    def graph(sel=slice(None)):
        node = cases.osc(ns, 'Sine', [hz[sel]])
        for k in range(2):
            r = np.random.default_rng(90 + k)
            lo, hi = r.uniform(500.0, 900.0, ch)[sel], r.uniform(2000.0, 6000.0, ch)[sel]
            wah = cases._wah(ns, [lo], [hi], [r.uniform(0.5, 3.0, ch)[sel]], [r.uniform(0.0, 1.0, ch)[sel]])
            node = cases._with_cutoff(ns, node, wah)
        return node

    got = {}
    for kernel in ('reg', 'scan'):
        c = engine.compile(graph(), ch, RATE)
        if kernel == 'scan':
            c.set_option('cascade_pipe', 0)
        got[kernel] = c.render_device(pos, frames)
        c.close()
    assert not bool(torch.equal(got['reg'], got['scan']))          # two different kernels did render
    diff = float((got['reg'] - got['scan']).abs().max())
    pick = np.sort(rng.choice(ch, 8, replace=False))
    want = np_oracle.GraphOracle(RATE).render(graph(pick), pos, frames, len(pick))
    err = max_abs_err(got['reg'][:, torch.from_numpy(pick).cuda()].cpu().numpy(), want)
    print(f'large modulated 2-section request: register kernel vs oracle {err:.3e}, vs the scan kernel {diff:.3e}')
    assert err <= 1e-4
    assert diff <= 2e-5


@pytest.mark.parametrize('fused', [1, 0])
def test_modulated_pan_matches_the_oracle_request_by_request(fused, ns, engine):
    """PanSum.pan driven by an LFO: sampled once per request at its first frame (block rate), so three consecutive requests
    pan the same voices differently; against the float64 oracle evaluating the same graph request by request -- fused with the
    voices (k_pan_weights re-derives the (L, R) weights per request) and on the materialised block (k_reduce reads the row)."""
    from signals_b200 import _lib
    from signals_b200.chain import ext
    n = 96
    prm = cases.instance_params(5, n)
    prm['gain'] = prm['gain'] * 4.0
    ps = cases.build_instances(ns, ext, prm)
    rng = np.random.default_rng(85)
    ps.pan = cases.sweep(ns, [rng.uniform(0.0, 0.3, n)], [rng.uniform(0.7, 1.0, n)], [rng.uniform(2.0, 9.0, n)], [rng.uniform(0.0, 1.0, n)])
    _lib.lib().sigb_set_default_option(b'fuse_reduce', fused)
    try:
        c = engine.compile(ps, 2, RATE)
        assert ('voices' in [l['kind'] for l in c.describe()['launches']]) == bool(fused)
        sizes = (4000, 1234, 9000)
        blocks, pos = [], 0
        for frames in sizes:
            blocks.append(c.render_device(pos, frames).cpu().numpy())
            pos += frames
        c.close()
    finally:
        _lib.lib().sigb_set_default_option(b'fuse_reduce', 1)
    # the oracle as one stream: filters carry their state across the requests (the plan's semantics), the pan is re-sampled
    # at each request's first frame
    orc = np_oracle.GraphOracle(RATE)
    flat = ps.inputs_by_port['input']
    x = orc.render(flat, 0, sum(sizes), n)                        # (frames, n): the voices before the PanSum, one request from 0
    want, pos = [], 0
    for frames in sizes:
        pan = orc.at_block_rate(ps, 'pan', pos, frames, n)
        want.append(np_oracle.pan_sum(x[pos:pos + frames], np.broadcast_to(pan, (1, n))))
        pos += frames
    err = max(max_abs_err(g, w) for g, w in zip(blocks, want))
    print(f'modulated pan over three requests (fused = {fused}): max-abs {err:.3e}')
    assert err <= 1e-6
    assert max_abs_err(want[0][:1000], np_oracle.pan_sum(x[:1000], np.broadcast_to(orc.at_block_rate(ps, 'pan', 4000, 10, n), (1, n)))) > 1e-4


@pytest.mark.parametrize('wave', ['Sine', 'Square', 'Sawtooth', 'Triangle'])
def test_wide_stateless_oscillator_chains(wave, ns, engine):
    """Oscillator -> gain on many channels (k_osc_fill: Q0.64 phase word per absolute 8-row block instead of float64 per sample):
    against the float64 oracle, bit for bit against itself however the stream is cut into requests (the blocks are aligned to
    the absolute sample index), against k_chain_seq on the same plan, at position 0 and far into the stream, with
    frequencies that put samples exactly on the discontinuities (6 kHz / 12 kHz at 48 kHz)."""
    rng = np.random.default_rng(86)
    ch = 200                                                  # ragged last tile
    hz = rng.uniform(27.5, 4186.0, ch)
    hz[:6] = [6000.0, 12000.0, 440.0, 439.99, 27.5, 8000.0]
    ph = rng.uniform(0.0, 1.0, ch)
    ph[:3] = [0.0, 0.25, 0.5]
    g = rng.uniform(0.05, 1.0, ch)
    node = cases.gain(ns, cases.osc(ns, wave, [hz], [ph]), [g])
    for pos in (0, 47999, 2 ** 31 + 3):
        frames = 5003
        c = engine.compile(node, ch, RATE)
        assert c.describe()['launches'][0]['sections'] == 0
        whole = c.render_device(pos, frames).cpu().numpy()
        cuts = [0, 1, 7, 8, 9, 48, 1000, 1001, 4097, frames]
        parts = np.concatenate([c.render_device(pos + a, b - a).cpu().numpy() for a, b in zip(cuts, cuts[1:])])
        c.set_option('osc_fill', 0)
        seq = c.render_device(pos, frames).cpu().numpy()
        c.close()
        want = np_oracle.GraphOracle(RATE).render(node, pos, frames, ch)
        err, err_seq = max_abs_err(whole, want), max_abs_err(seq, want)
        print(f'{wave} x {ch} channels at position {pos}: k_osc_fill max-abs {err:.3e} (k_chain_seq {err_seq:.3e})')
        assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32)), (wave, pos)
        assert not np.array_equal(whole, seq) or wave == 'Square'          # two different kernels did render
        # 1e-6 for oscillators; far into the stream the reference's own float64 phase is rounded to ~4e-7 rad at 12 kHz (cases.py)
        assert err <= (1e-6 if pos < 2 ** 31 else 2e-6), (wave, pos, err)


@pytest.mark.parametrize('wave,nsec', [('Sine', 0), ('Sine', 1), ('Sine', 2), ('Sine', 3), ('Square', 0), ('Square', 1), ('Sawtooth', 0),
                                       ('Sawtooth', 2), ('Triangle', 1), ('Triangle', 3)])
def test_vibrato_keeps_the_phase_word_fast_paths(wave, nsec, ns, engine):
    """An oscillator whose hertz AND phase are driven by LFOs on many channels: the rows are constant within a request, so
    k_osc_tables re-derives the exact Q0.64 phase tables per request and the chain runs on k_osc_fill / k_chain_scan3 (Sine) /
    k_osc_delta like a constant oscillator (the discontinuous waveforms with a guard band from the thread's own channels) --
    three consecutive requests (the parameters jump between them, as in the reference) against the float64 oracle and against
    the float64-per-sample kernel (force_seq)."""
    ch = 192
    rng = np.random.default_rng(87 + nsec)
    hz = rng.uniform(55.0, 3000.0, ch)
    o = getattr(ns, wave)()
    o.hertz = cases.sweep(ns, [hz * 0.94], [hz * 1.06], [rng.uniform(3.0, 7.0, ch)], [rng.uniform(0, 1, ch)])
    o.phase = cases.gain(ns, cases.osc(ns, 'Sine', [rng.uniform(0.5, 2.0, ch)], [rng.uniform(0, 1, ch)]), [np.full(ch, 0.1)])
    node = o
    cut = np.exp(rng.uniform(np.log(700.0), np.log(8000.0), (max(nsec, 1), ch)))
    for s in range(nsec):
        node = cases.lowpass(ns, node, [cut[s]])
    node = cases.gain(ns, node, [rng.uniform(0.1, 1.0, ch)])
    sizes = (20000, 4099, 24000)
    got = {}
    for mode in ('fast', 'seq'):
        c = engine.compile(node, ch, RATE)
        if mode == 'seq':
            c.set_option('force_seq', 1)
        blocks, pos = [], 0
        for frames in sizes:
            blocks.append(c.render_device(pos, frames).cpu().numpy())
            pos += frames
        c.close()
        got[mode] = np.concatenate(blocks)
    assert not np.array_equal(got['fast'], got['seq']) or (wave == 'Square' and nsec == 0)     # two different kernels did render
    tol = 1e-6 if nsec == 0 else 1e-4
    assert max_abs_err(got['fast'], got['seq']) <= (2e-6 if nsec == 0 else 2e-5)
    if nsec == 0:            # (filters: the plan carries their state across the requests, the oracle's requests restart; compare with seq only)
        orc = np_oracle.GraphOracle(RATE)
        want, pos = [], 0
        for frames in sizes:
            want.append(orc.render(node, pos, frames, ch))
            pos += frames
        err = max_abs_err(got['fast'], np.concatenate(want))
        print(f'vibrato {wave} x {ch} channels, three requests: max-abs {err:.3e}')
        assert err <= tol


@pytest.mark.parametrize('wave', ['Square', 'Sawtooth', 'Triangle'])
def test_one_section_behind_a_discontinuous_waveform_goes_register_resident(wave, ns, engine):
    """Square / Sawtooth / Triangle -> ONE filter -> gain on many channels (lowpass_test.sigs' shape at scale): the scan kernels
    evaluate those waveforms in float64 per sample, so the chain takes k_osc_delta when its time pieces fill the machine --
    against the oracle on picked voices and against the scan kernel (osc_reg = 0) on the same plan."""
    torch = pytest.importorskip('torch')
    ch, frames = 4096, 96000
    hz, ph, cut, g = cases.voice_params(88, ch)
    cut = np.maximum(cut, 800.0)
    node = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, wave, [hz], [ph]), [cut]), [g])
    got = {}
    for kernel in ('reg', 'scan'):
        c = engine.compile(node, ch, RATE)
        if kernel == 'scan':
            c.set_option('osc_reg', 0)
        first = c.render_device(0, frames)
        second = c.render_device(frames, 1000)
        got[kernel] = torch.cat([first, second])
        c.close()
    assert not bool(torch.equal(got['reg'], got['scan']))
    diff = float((got['reg'] - got['scan']).abs().max())
    pick = np.sort(np.random.default_rng(3).choice(ch, 8, replace=False))
    sub = cases.gain(ns, cases.lowpass(ns, cases.osc(ns, wave, [hz[pick]], [ph[pick]]), [cut[pick]]), [g[pick]])
    want = np_oracle.GraphOracle(RATE).render(sub, 0, frames + 1000, len(pick))
    err = max_abs_err(got['reg'][:, torch.from_numpy(pick).cuda()].cpu().numpy(), want)
    print(f'{wave} -> LowPass -> Gain x {ch}: register kernel max-abs {err:.3e}; vs the scan kernel {diff:.3e}')
    assert err <= 1e-4
    assert diff <= 2e-5


@pytest.mark.parametrize('op,wa,wb,side', [('Mix', 'Sine', 'Sine', 0), ('Mix', 'Sawtooth', 'Sine', 1), ('RingMod', 'Sine', 'Square', 0),
                                           ('RingMod', 'Triangle', 'Sawtooth', 1), ('Mix', 'Square', 'Triangle', 0)])
def test_wide_mix_and_ringmod_of_two_oscillators(op, wa, wb, side, ns, engine):
    """Mix / RingMod of two oscillator chains on many channels: fused into k_osc_fill (both oscillators from their Q0.64 phase
    words, in registers) -- against the float64 oracle, bit for bit however the stream is cut, against k_chain_seq's fused
    epilogue on the same plan."""
    rng = np.random.default_rng(89)
    ch = 200
    a = cases.gain(ns, cases.osc(ns, wa, [rng.uniform(27.5, 4186.0, ch)], [rng.uniform(0, 1, ch)]), [rng.uniform(0.1, 1.0, ch)])
    b = cases.gain(ns, cases.osc(ns, wb, [rng.uniform(27.5, 4186.0, ch)], [rng.uniform(0, 1, ch)]), [rng.uniform(0.1, 1.0, ch)])
    node = getattr(ns, op)()
    node.left, node.right = (b, a) if side else (a, b)
    if op == 'Mix':
        node.mix = cases.fixed(ns, [rng.uniform(0.0, 1.0, ch)])
    for pos in (0, 123457):
        frames = 4099
        c = engine.compile(node, ch, RATE)
        assert [l['kind'] for l in c.describe()['launches']] == ['chain']            # one fused launch
        whole = c.render_device(pos, frames).cpu().numpy()
        cuts = [0, 1, 7, 8, 9, 1000, 1001, frames]
        parts = np.concatenate([c.render_device(pos + x, y - x).cpu().numpy() for x, y in zip(cuts, cuts[1:])])
        c.set_option('osc_fill', 0)
        seq = c.render_device(pos, frames).cpu().numpy()
        c.close()
        want = np_oracle.GraphOracle(RATE).render(node, pos, frames, ch)
        err = max_abs_err(whole, want)
        print(f'{op}({wa}, {wb}) x {ch} at position {pos}: k_osc_fill max-abs {err:.3e} (k_chain_seq {max_abs_err(seq, want):.3e})')
        assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))
        assert not np.array_equal(whole, seq)
        assert err <= 1e-6


@pytest.mark.parametrize('nsec', [0, 1, 2])
def test_tremolo_rides_on_the_chain_gain(nsec, ns, engine):
    """A Gain driven by an LFO at the END of a chain (tremolo) on many channels: folded into the chain's gain table per request
    (k_gain_rows) -- one launch per request's block, no pointwise pass -- against the float64 oracle over three requests
    (oscillator only) and against the unfused plan (fuse_pointwise = 0: chain + k_ewise)."""
    from signals_b200 import _lib
    ch = 192
    rng = np.random.default_rng(90 + nsec)
    node = cases.osc(ns, 'Sine', [rng.uniform(55.0, 3000.0, ch)], [rng.uniform(0, 1, ch)])
    cut = np.exp(rng.uniform(np.log(700.0), np.log(8000.0), (max(nsec, 1), ch)))
    for s in range(nsec):
        node = cases.lowpass(ns, node, [cut[s]])
    trem = ns.Gain()
    trem.left = node
    trem.right = cases.sweep(ns, [np.full(ch, 0.2)], [np.full(ch, 1.0)], [rng.uniform(3.0, 9.0, ch)], [rng.uniform(0, 1, ch)])
    sizes = (20000, 4099, 24000)
    got = {}
    for fused in (1, 0):
        _lib.lib().sigb_set_default_option(b'fuse_pointwise', fused)
        try:
            c = engine.compile(trem, ch, RATE)
            kinds = [l['kind'] for l in c.describe()['launches']]
            assert kinds == (['chain'] if fused else ['chain', 'ewise']), kinds
            blocks, pos = [], 0
            for frames in sizes:
                blocks.append(c.render_device(pos, frames).cpu().numpy())
                pos += frames
            c.close()
        finally:
            _lib.lib().sigb_set_default_option(b'fuse_pointwise', 1)
        got[fused] = np.concatenate(blocks)
    assert max_abs_err(got[1], got[0]) <= (1e-6 if nsec == 0 else 2e-5)
    if nsec == 0:
        orc = np_oracle.GraphOracle(RATE)
        want, pos = [], 0
        for frames in sizes:
            want.append(orc.render(trem, pos, frames, ch))
            pos += frames
        err = max_abs_err(got[1], np.concatenate(want))
        print(f'tremolo on {ch} sines, three requests: max-abs {err:.3e}')
        assert err <= 1e-6
