"""The N>1 path on CPU: two gloo ranks shard the channels of one batch, each runs its
range (through the oracle here — no GPU), and the gathered result equals the single-rank
run; timings reduce as max, units as sum.  This is the host logic bench.py uses under
torch.distributed.run on the GPU box (NCCL there)."""
import os
import socket

import numpy as np
import pytest

from jsdrcuda import sharding


def test_partition_is_contiguous_balanced_and_complete():
    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [sharding.partition(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.partition(8, 2, 2)


def test_channel_tuning_is_independent_of_the_sharding():
    whole = sharding.channel_tuning(0, 10)
    parts = np.concatenate([sharding.channel_tuning(*sharding.partition(10, 3, r)) for r in range(3)])
    assert np.array_equal(whole, parts)
    assert whole.min() >= 2000.0 and whole.max() <= 90000.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nchan, S, rate, out_dir):
    import torch
    import torch.distributed as dist
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.partition(nchan, world, rank)
    tun = sharding.channel_tuning(first, count)
    rng = np.random.Generator(np.random.PCG64(123))
    raw = rng.integers(-20000, 20000, (nchan, 2 * S)).astype(np.int16)      # same batch on every rank
    ds, _ = O.baseline_mixdecim_s16(raw[first:first + count], count, rate, tun, None, 1)
    ms, units = sharding.reduce_timing(dist, torch.device("cpu"), [10.0 + rank, 5.0 - rank], count * S)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, ds))
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])])
        np.save(os.path.join(out_dir, "ds.npy"), full)
        np.save(os.path.join(out_dir, "red.npy"), np.array(ms + [units]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp
    import oracle as O
    nchan, S, rate, world = 6, 2400, 96000, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nchan, S, rate, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "ds.npy")
    red = np.load(tmp_path / "red.npy")
    rng = np.random.Generator(np.random.PCG64(123))
    raw = rng.integers(-20000, 20000, (nchan, 2 * S)).astype(np.int16)
    ref, _ = O.baseline_mixdecim_s16(raw, nchan, rate, sharding.channel_tuning(0, nchan), None, 1)
    assert np.array_equal(full, ref)                       # sharding by channel changes nothing
    assert red[0] == 11.0 and red[1] == 5.0                # max over ranks
    assert red[2] == nchan * S                             # units summed
