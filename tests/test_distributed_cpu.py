"""world_size-2 gloo tests of the N>1 host logic (sharding + final metric reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brever_b200.distributed import global_mean, shard, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_bounds(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            for (a, b), (c, d) in zip(pieces, pieces[1:]):
                assert b == c
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        metric = torch.randn(n_items, generator=g)       # same on every rank
        mine = shard(metric)                             # this rank's utterances
        mean = global_mean(mine)
        torch.save({'mean': mean, 'n': mine.numel()}, os.path.join(result_dir, f'{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_items', [64, 7])
def test_two_rank_metric_reduce(tmp_path, n_items):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    expect = torch.randn(n_items, generator=g).mean()
    out = [torch.load(tmp_path / f'{r}.pt') for r in range(world)]
    assert sum(o['n'] for o in out) == n_items
    for o in out:
        assert torch.allclose(o['mean'], expect, atol=1e-6)


def test_single_process_is_identity():
    x = torch.arange(10.)
    assert torch.equal(shard(x, 0, 1), x)
    assert torch.allclose(global_mean(x), x.mean())
