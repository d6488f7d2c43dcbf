"""world_size-2 gloo test of the N>1 host logic (batch sharding + token gather) on CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from p3tok import shard


def test_shard_slice_is_a_partition():
    for B in (0, 1, 7, 128, 4096):
        for W in (1, 2, 3, 8):
            spans = [shard.shard_slice(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_slice(4, 2, 2)


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(B * 3 * 4, dtype=torch.float32).view(B, 3, 4)
        mine = shard.shard_batch(full)                      # each rank "tokenizes" its slice
        lo, hi = shard.shard_slice(B, rank, world)
        assert mine.shape[0] == hi - lo
        got = shard.gather_tokens(mine * 2.0)
        q.put((rank, bool(torch.equal(got, full * 2.0))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_gather_tokens_world2_gloo(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _syncbn_worker(rank, world, port, q):
    """The SyncBN host logic of the training path (p3tok.train): every rank contributes (sum, sum of squares, rows) of its
    shard; after the all-reduce the combined statistics equal those of the whole batch (the kernels themselves are CUDA-only;
    here the per-shard sums come from torch)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from p3tok import train
        torch.manual_seed(11)
        full = torch.randn(96, 5, dtype=torch.float64) * 2 + 0.5
        lo, hi = shard.shard_slice(96, rank, world)
        x = full[lo:hi]
        packed = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(hi - lo)], dtype=torch.float64)])
        train._all_reduce(packed, True)
        mean, rstd, var_unb = train.combine_stats(packed[:5], packed[5:10], float(packed[10]), 1e-5)
        ok = (torch.allclose(mean, full.mean(0)) and torch.allclose(var_unb, full.var(0, unbiased=True))
              and torch.allclose(rstd, 1.0 / torch.sqrt(full.var(0, unbiased=False) + 1e-5)) and float(packed[10]) == 96.0)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_syncbn_statistics_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, 29655, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
