"""CPU, world_size 2 over gloo: the N>1 path of signals_b200.shard (voice partition + the single reduce
of the stereo mix).  The partial mixes are made by the numpy oracle here (no GPU in this suite); on a
GPU box the same reduce_mix call runs on NCCL with the partials rendered by k_voices."""
import os
import socket

import numpy as np
import pytest

from oracle import cases, np_oracle
from signals_b200 import shard

RATE = 48000


def test_shard_indices_partition_the_bank():
    for n, world in [(10, 1), (10, 3), (1 << 20, 8), (5, 8)]:
        parts = [shard.shard_indices(n, r, world) for r in range(world)]
        allv = np.sort(np.concatenate(parts))
        assert np.array_equal(allv, np.arange(n))
        assert [len(p) for p in parts] == shard.shard_counts(n, world)
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def test_instance_params_shards_are_slices_of_one_bank():
    full = cases.instance_params(5, 1000)
    for world in (2, 4, 8):
        for r in range(world):
            mine = cases.instance_params(5, 1000, r, world)
            idx = shard.shard_indices(1000, r, world)
            for k in ('wave', 'filt', 'hertz', 'phase', 'cutoff', 'gain', 'pan'):
                assert np.array_equal(mine[k], full[k][idx]), k


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, frames, dst, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        prm = cases.instance_params(5, n, rank, world)
        partial = torch.from_numpy(np_oracle.render_instances(prm, 0, frames, RATE).astype(np.float32))
        mine = partial.clone()
        mix = shard.reduce_mix(partial, dst=dst)
        q.put((rank, mix.numpy().copy(), mine.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('dst', [0, None])
def test_reduce_mix_world2_equals_unsharded_oracle(dst):
    import torch.multiprocessing as mp
    n, frames, world = 48, 2400, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, frames, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, mix, mine = q.get(timeout=120)
        got[rank] = (mix, mine)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np_oracle.render_instances(cases.instance_params(5, n), 0, frames, RATE)
    assert np.abs(got[0][0] - want).max() <= 1e-6
    if dst is None:                                   # all-reduce: every rank holds the mix
        assert np.array_equal(got[1][0], got[0][0])
    # the reduce is exactly the sum of the two partials (float32 add, one rounding)
    assert np.array_equal(got[0][0], got[0][1] + got[1][1])


def test_reduce_mix_without_process_group_is_identity():
    import torch
    t = torch.arange(6, dtype=torch.float32).reshape(3, 2)
    assert shard.reduce_mix(t) is t
