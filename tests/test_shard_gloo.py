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
