"""SURVEY 8f rows built after the bar: the `.sigs` patch loader, plugin discovery by dotted name, the
headless sounddevice shim and the SinkDevice adapter (the caller of the hot path, chain/dev.py:167-179).
CPU tests cover parsing / lowering / plumbing; the `gpu` tests render through the C ABI."""
import io
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT, load_golden, max_abs_err
from oracle import np_oracle
from signals_b200 import _lib, sigs, sounddevice_shim
from signals_b200.chain import discovery, dev, fx, osc, vis

RATE = 48000
SCRIPT = os.path.join(ROOT, 'tests', 'scripts', 'sine_sink.py')


def test_parse_value_follows_the_reference_syntax():
    assert sigs.parse_value('1') == 1 and sigs.parse_value('true') is True and sigs.parse_value('-1.0') == -1.0
    v = sigs.parse_value('[[1, 2, 3]]')
    assert isinstance(v, np.ndarray) and v.shape == (1, 3)
    assert sigs.parse_value('/tmp/lowpass_test.wav') == '/tmp/lowpass_test.wav'       # not JSON -> raw string


def test_load_signal_maps_reference_names_onto_the_mirror():
    assert discovery.load_signal('signals.chain.osc.Sine') is osc.Sine
    assert discovery.load_signal('signals_b200.chain.fx.LowPass') is fx.LowPass
    assert discovery.load_signal('signals.chain.vis.Wave') is vis.Wave
    with pytest.raises(discovery.BadPath):
        discovery.load_signal('signals.chain.osc.Nope')
    with pytest.raises(discovery.InvalidObject):
        discovery.load_signal('signals.chain.osc.Osc')          # abstract
    with pytest.raises(discovery.BadSyntax):
        discovery.load_signal('not a name')


@pytest.mark.parametrize('name,root_kind,launch', [
    ('lowpass_test', 'Wave', dict(kind='chain', source='osc', wave='triangle', sections=1, gain=True)),
    ('vis_test', 'Wave', dict(kind='chain', source='osc', wave='sine', sections=0, gain=False)),
])
def test_reference_patch_files_load_and_lower(name, root_kind, launch, engine):
    """The reference's two patch fixtures (src/signals/*.sigs, copied as data to tests/golden/) replay
    headlessly; taps (Wave, FileWriter) lower to their inputs, so each patch is ONE fused chain launch."""
    patch = sigs.load(os.path.join(GOLDEN_DIR, name + '.sigs'))
    sink, emitter = patch.root()
    assert type(emitter).__name__ == root_kind and sink.get_state().channels == 1
    (got,) = engine.compile(emitter, 1, RATE).describe()['launches']
    assert {k: got[k] for k in launch} == launch


def test_patch_commands_and_errors():
    patch = sigs.loads('''
        sink 9a default channels=2
        + 1a signals.chain.fixed.Fixed value=[[220,330]]   # comment
        + 2a signals.chain.osc.Square
        > 1a 2a.hertz
        > 2a 9a.input
    ''')
    sink, emitter = patch.root()
    assert sink.get_state().channels == 2 and emitter.channels == 2
    patch.execute('* 1a value=[[1,2]]')
    assert patch.nodes['1a'].get_state().value.tolist() == [[1, 2]]
    patch.execute('>/ 9a.input')
    with pytest.raises(sigs.PatchError):
        patch.root()
    patch.execute('- 2a')
    assert '2a' not in patch.nodes
    for bad in ('+ 1a signals.chain.osc.Sine', '> 1a 9a.nope', 'play', '* 3c x=1', '+ zz signals.chain.osc.Sine'):
        with pytest.raises((sigs.PatchError, discovery.DiscoveryError)):
            patch.execute(bad)
    with pytest.raises(sigs.PatchError, match='line 2'):
        sigs.loads('sink 1a default\n> 5a 1a.input\n')


def test_shim_drives_the_callback_and_records_blocks():
    calls = []

    def callback(outdata, frames, time, status):
        calls.append(frames)
        outdata[:] = len(calls)
        if len(calls) == 3:
            raise sounddevice_shim.CallbackStop

    sounddevice_shim.install(blocks=5, blocksize=64, force=False)
    stream = sounddevice_shim.OutputStream(channels=2, callback=callback)
    stream.start()
    assert calls == [64, 64, 64] and not stream.active
    audio = stream.audio()
    assert audio.shape == (192, 2) and audio.dtype == np.float32 and audio[-1, 0] == 3.0
    assert sounddevice_shim.query_devices()[0]['name'] == 'default'


def test_sink_device_mirrors_the_reference_surface():
    rack = discovery.Rack()
    rack.scan()
    sink = dev.SinkDevice(rack.get_sink('default'))
    assert sink.port_names() == ['input'] and not sink.is_open and sink.frame_position == 0
    with pytest.raises(dev.BadPlaybackState):
        sink.close()
    with pytest.raises(discovery.BadDeviceName):
        rack.get_sink('nope')
    sink.open()
    with pytest.raises(dev.BadPlaybackState):
        sink.open()
    with pytest.raises(dev.BadPlaybackState):
        sink.stop()
    sink.destroy()
    assert not sink.is_open


def test_script_runs_unchanged_up_to_the_render(capsys):
    """Without a GPU the script still builds its graph and starts the sink; the first callback fails loudly
    (no CPU fallback) and the stream stops, exactly like the reference's except-branch (dev.py:174-176)."""
    from signals_b200 import run_script
    if _lib.lib().sigb_device_count() > 0:
        pytest.skip('GPU present: covered by the gpu test')
    streams = run_script.run(SCRIPT, blocks=4, blocksize=128)
    assert len(streams) >= 1 and len(streams[-1].recorded) == 1
    assert 'no CPU fallback' in capsys.readouterr().err


def test_reference_scripts_run_unchanged_when_present(capsys):
    """scripts/edited_sine.py and scripts/example_sine.py of the reference (build container only)."""
    from signals_b200 import run_script
    ref = '/root/reference/scripts'
    if not os.path.isdir(ref):
        pytest.skip('reference sources only exist in the build container')
    n0 = len(sounddevice_shim.streams)
    run_script.run(os.path.join(ref, 'edited_sine.py'), blocks=2, blocksize=64)
    assert len(sounddevice_shim.streams) == n0 + 1
    run_script.run(os.path.join(ref, 'example_sine.py'), blocks=2, blocksize=64, stdin_text='\n')
    audio = sounddevice_shim.streams[-1].audio()      # the upstream example computes in its own callback (numpy)
    assert max_abs_err(audio, np_oracle.example_sine_block(0, 128, 48000.0)) <= 1e-6


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_patch_files_render_like_the_reference():
    got = sigs.load(os.path.join(GOLDEN_DIR, 'vis_test.sigs')).render(0, 4800)
    assert max_abs_err(got, load_golden('vis_test_sigs')) <= 1e-6
    # lowpass_test.sigs: the sink hangs off Wave <- FileWriter <- LowPass (column 0 of the golden Merge render)
    got = sigs.load(os.path.join(GOLDEN_DIR, 'lowpass_test.sigs')).render(0, 48000)
    assert got.shape == (48000, 1)
    assert max_abs_err(got[:, 0], load_golden('lowpass_test_sigs')[:, 0]) <= 1e-4


@pytest.mark.gpu
def test_sink_device_blockwise_equals_single_request(ns):
    """The realtime adapter: 512-frame callbacks with carried filter state == one request (the reference
    itself cannot do this: it restarts every block from zero state)."""
    from oracle import cases
    sounddevice_shim.install(blocks=20, blocksize=512)
    rack = discovery.Rack()
    rack.scan()
    sink = dev.SinkDevice(rack.get_sink('default'))
    st = sink.get_state()
    import attr
    sink.set_state(attr.evolve(st, channels=8))
    sink.input = cases.CASES_BY_NAME['lowpass_c2_8v'].build(ns)
    sink.start()
    audio = sink._stream.audio()
    sink.destroy()
    assert audio.shape == (20 * 512, 8) and sink.frame_position == 20 * 512
    assert max_abs_err(audio, load_golden('lowpass_c2_8v')[:20 * 512]) <= 1e-4


@pytest.mark.gpu
def test_script_renders_through_the_sink_device():
    from signals_b200 import run_script
    streams = run_script.run(SCRIPT, blocks=10, blocksize=480)
    audio = streams[-1].audio()
    assert audio.shape == (4800, 1)
    assert max_abs_err(audio, np_oracle.example_sine_block(0, 4800, 48000.0)) <= 1e-6


# ------------------------------------------------------------------------------------------------
# file nodes
# ------------------------------------------------------------------------------------------------

def test_wav_codec_round_trip(tmp_path):
    from signals_b200 import wavio
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (1000, 3)).astype(np.float32)
    path = str(tmp_path / 'a.wav')
    w = wavio.WavWriter(path, 48000, 3)
    w.write(0, x[:400])
    w.write(400, x[400:])
    w.write(100, x[100:200])                      # positioned rewrite, like SoundFile.seek + write (files.py:54-56)
    w.close()
    y, rate = wavio.read(path)
    assert rate == 48000 and y.dtype == np.float32 and np.array_equal(x, y)
    # a 16-bit PCM file as other tools write them
    import struct
    pcm = (x[:, :1] * 32767).astype('<i2')
    with open(tmp_path / 'b.wav', 'wb') as f:
        f.write(b'RIFF' + struct.pack('<I', 36 + pcm.nbytes) + b'WAVEfmt ' + struct.pack('<IHHIIHH', 16, 1, 1, 44100, 88200, 2, 16))
        f.write(b'data' + struct.pack('<I', pcm.nbytes) + pcm.tobytes())
    z, rate = wavio.read(str(tmp_path / 'b.wav'))
    assert rate == 44100 and z.shape == (1000, 1) and np.abs(z - pcm / 32768.0).max() == 0.0


def test_file_reader_lowers_to_an_hbm_buffer(tmp_path, engine):
    from signals_b200 import wavio
    from signals_b200.chain import files
    x = np.linspace(-1, 1, 600, dtype=np.float32).reshape(300, 2)
    w = wavio.WavWriter(str(tmp_path / 'in.wav'), 48000, 2)
    w.write(0, x)
    w.close()
    r = files.FileReader()
    r.get_state().path = str(tmp_path / 'in.wav')
    assert r.channels == 2 and r.file_rate == 48000 and np.array_equal(r.samples, x)
    f = fx.HighPass()
    f.input = r
    from oracle import cases
    f.cutoff = cases.fixed(cases.b200_namespace(), [[500.0, 900.0]])
    d = engine.compile(f, 2, 48000).describe()
    assert d['launches'] == [d['launches'][0]] and d['launches'][0]['source'] == 'block' and d['launches'][0]['sections'] == 1
    assert discovery.load_signal('signals.chain.files.FileReader') is files.FileReader


@pytest.mark.gpu
def test_file_nodes_on_the_gpu_path(tmp_path):
    """FileReader -> LowPass -> FileWriter -> Wave -> sink, written as a patch: the sink block equals the oracle's
    filter of the file, the FileWriter's WAV holds exactly the block that passed through it, and the Wave tap
    queued it -- blockwise, with the filter state carried across the requests."""
    from signals_b200 import engine, wavio
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (6000, 2)).astype(np.float32)
    w = wavio.WavWriter(str(tmp_path / 'in.wav'), 48000, 2)
    w.write(0, x)
    w.close()
    patch = sigs.loads(f"""
        sink 9a default channels=2
        + 1a signals.chain.files.FileReader path={tmp_path / 'in.wav'}
        + 1b signals.chain.fixed.Fixed value=[[1200,3000]]
        + 2a signals.chain.fx.LowPass
        + 3a signals.chain.files.FileWriter path={tmp_path / 'out.wav'}
        + 4a signals.chain.vis.Wave
        > 1a 2a.input
        > 1b 2a.cutoff
        > 2a 3a.input
        > 3a 4a.input
        > 4a 9a.input
    """)
    engine.default_engine().clear()
    blocks = [patch.render(p, 1500, taps=True) for p in range(0, 6000, 1500)]
    got = np.concatenate(blocks)
    want, _ = np_oracle.render_cascade(x.astype(np.float64), np.array([[1200.0, 3000.0]]), RATE)
    assert max_abs_err(got, want) <= 1e-4
    patch.nodes['3a'].destroy()                                     # closes the file
    written, rate = wavio.read(str(tmp_path / 'out.wav'))
    assert rate == RATE and written.shape == (6000, 2)
    # the taps re-render their input per request from a fresh plan position sequence: same carried state, same bits
    assert max_abs_err(written, got) <= 1e-6
    q = patch.nodes['4a'].q
    assert q.qsize() == 4 and q.get().shape == (1500, 2)
