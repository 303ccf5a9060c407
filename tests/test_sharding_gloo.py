"""The N>1 host logic on CPU with gloo, world_size 2: shard ranges tile [0, N) exactly and the
weight-statistics all-reduce combines max / sum entries correctly."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phasespace_b200 import sharding


def test_shard_ranges_partition_the_events():
    for n in (0, 1, 7, 8, 9, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            pos = 0
            for rank in range(world):
                first, count = sharding.shard_range(n, rank, world)
                assert first == pos and count >= 0
                pos += count
            assert pos == n
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, n_events, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = sharding.shard_range(n_events, *sharding._rank_world())
        # per-rank statistics as the kernel epilogue would produce them: [max w, sum w, sum w^2, max w_max]
        w = torch.arange(first, first + count, dtype=torch.float64) / n_events
        stats = torch.stack([w.max(), w.sum(), (w * w).sum(), torch.tensor(10.0 + rank, dtype=torch.float64)])
        red = sharding.allreduce_stats(stats)
        total = sharding.allreduce_count(count)
        hist = sharding.allreduce_sum(torch.bincount((w * 10).long(), minlength=10).to(torch.float64))
        # the global maximum weight of a sharded GenMultiDecay (one scalar MAX)
        w_max = sharding.allreduce_max(torch.tensor(3.0 + 2.0 * rank, dtype=torch.float64), world=world)
        out[rank] = (first, count, red.tolist(), total, hist.tolist(), float(w_max))
    finally:
        dist.destroy_process_group()


def test_two_rank_statistics_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_events, world = 1001, 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_events, out), nprocs=world, join=True)
    w = torch.arange(n_events, dtype=torch.float64) / n_events
    expect = [w.max().item(), w.sum().item(), (w * w).sum().item(), 11.0]
    assert out[0][0] == 0 and out[1][0] == out[0][1] and out[0][1] + out[1][1] == n_events
    for rank in range(world):
        assert out[rank][3] == n_events
        assert out[rank][5] == 5.0
        assert out[rank][4] == torch.bincount((w * 10).long(), minlength=10).to(torch.float64).tolist()
        for got, want in zip(out[rank][2], expect):
            assert got == pytest.approx(want, rel=1e-12)


def test_single_process_is_identity():
    stats = torch.tensor([0.5, 10.0, 3.0, 7.0], dtype=torch.float64)
    assert torch.equal(sharding.allreduce_stats(stats), stats)
    assert sharding.allreduce_count(5) == 5
    assert float(sharding.allreduce_max(torch.tensor(2.5, dtype=torch.float64))) == 2.5
    with pytest.raises(RuntimeError):  # ranks without a process group cannot agree on a maximum
        sharding.allreduce_max(torch.tensor(2.5, dtype=torch.float64), world=2)


def test_multidecay_mode_counts_depend_on_the_seed_only():
    """Every rank of a sharded GenMultiDecay computes the split of the events over the decay modes on its own host: the
    same (seed, offset) must give the same counts, summing to n_events (genmultidecay.py:165-183)."""
    from phasespace_b200 import GenParticle
    from phasespace_b200.fromdecay import GenMultiDecay
    from phasespace_b200.random import Generator

    modes = [(p, GenParticle(f"T{k}", 1000.0).set_children(GenParticle(f"a{k}", 100.0), GenParticle(f"b{k}", 200.0)))
             for k, p in enumerate((0.5, 0.3, 0.15, 0.05))]
    gmd = GenMultiDecay(modes)
    a, b = gmd.mode_counts(100_000, Generator(7)), gmd.mode_counts(100_000, Generator(7))
    assert a.sum() == 100_000 and (a == b).all() and (a > 0).all()
    assert (gmd.mode_counts(100_000, Generator(8)) != a).any()
    assert abs(a[0] / 100_000 - 0.5) < 0.01
