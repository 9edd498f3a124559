"""TEST INFRASTRUCTURE ONLY -- float64 oracle renders of the BASELINE configs at FULL width, reduced to small fixtures
(run once in the build container: a few minutes on 8 cores; the fixtures are committed under tests/golden/):

  full_c2_grid.npz   C2, 4,096 voices x 480,000 frames: sums over cells of 1,000 rows x 64 voices -> (480, 64)
  full_c4_grid.npz   C4, 16,384 channels x 8 sections, first 48,000 frames of a numpy-seeded noise block: cells of
                     1,000 rows x 64 channels -> (48, 256)
  full_c5_mix.npz    C5, 1,048,576 instances, first 4,800 frames of the stereo mix -> (4800, 2)

A cell sum moves by hundreds when one 64-channel x 16-row tile of the block is wrong, and by < 0.1 under the 1e-6 / 1e-4
per-sample budgets, so the GPU tests that compare against these grids check EVERY tile of the full-size blocks, not a sample of
voices.  The oracle functions are the ones pinned against the reference's goldens (tests/test_oracle.py).

    python -m oracle.make_fullsize [c2] [c4] [c5]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

from oracle import cases, np_oracle

RATE = 48000
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
CELL_ROWS, CELL_CH = 1000, 64


def c4_noise(frames: int, ch: int) -> np.ndarray:
    """The C4 test input: U(-1, 1) float32 from a numpy generator, so that the CPU oracle and the GPU test read the same block."""
    return np.random.default_rng(44).random((frames, ch), dtype=np.float32) * 2.0 - 1.0


def _c2_job(i):
    hertz, phase, cutoff, g = cases.voice_params(2, 4096)
    sl = slice(i * CELL_CH, (i + 1) * CELL_CH)
    y = np_oracle.render_voice_chain(0, 10 * RATE, RATE, hertz[sl], phase[sl], cutoff[sl], g[sl])
    return y.reshape(-1, CELL_ROWS, CELL_CH).sum(axis=(1, 2))


def _c4_job(i):
    rng = np.random.default_rng(4)
    cut = np.exp(rng.uniform(np.log(200.0), np.log(8000.0), (8, 16384)))
    sl = slice(i * CELL_CH, (i + 1) * CELL_CH)
    x = _C4_X[:, sl].astype(np.float64)
    y, _ = np_oracle.render_cascade(x, cut[:, sl], RATE)
    return y.reshape(-1, CELL_ROWS, CELL_CH).sum(axis=(1, 2))


def _c5_job(i):
    prm = cases.instance_params(5, 1 << 20, i, 512)         # a 1/512 slice of the bank (2048 instances)
    return np_oracle.render_instances(prm, 0, 4800, RATE)


_C4_X = None


def main(which):
    global _C4_X
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    workers = os.cpu_count() or 1
    if 'c2' in which:
        t0 = time.time()
        with mp.get_context('fork').Pool(workers) as pool:
            cols = pool.map(_c2_job, range(4096 // CELL_CH), chunksize=1)
        grid = np.stack(cols, axis=1)
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'full_c2_grid.npz'), grid=grid)
        print('c2 grid', grid.shape, 'abs max', np.abs(grid).max(), '%.0f s' % (time.time() - t0))
    if 'c4' in which:
        t0 = time.time()
        _C4_X = c4_noise(48000, 16384)
        with mp.get_context('fork').Pool(workers) as pool:
            cols = pool.map(_c4_job, range(16384 // CELL_CH), chunksize=1)
        grid = np.stack(cols, axis=1)
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'full_c4_grid.npz'), grid=grid)
        print('c4 grid', grid.shape, 'abs max', np.abs(grid).max(), '%.0f s' % (time.time() - t0))
    if 'c5' in which:
        t0 = time.time()
        with mp.get_context('fork').Pool(workers) as pool:
            parts = pool.map(_c5_job, range(512), chunksize=4)
        mix = np.sum(parts, axis=0)
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'full_c5_mix.npz'), mix=mix)
        print('c5 mix', mix.shape, 'abs max', np.abs(mix).max(), '%.0f s' % (time.time() - t0))


if __name__ == '__main__':
    main(sys.argv[1:] or ['c2', 'c4', 'c5'])
