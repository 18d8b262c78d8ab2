"""The N>1 path on CPU: chain sharding and the sample all-gather over a world_size-2 gloo group (no GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_every_chain_once(pkg):
    from bayes_js_b200.parallel import shard_bounds
    for n, ws in ((1 << 20, 8), (1000, 7), (5, 8), (64, 2), (1, 1)):
        blocks = [shard_bounds(n, r, ws) for r in range(ws)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
            assert f1 == f0 + c0                       # contiguous, in rank order
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, n_chains, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    graft.load_package()
    from bayes_js_b200.parallel import all_gather_chain_axis, gather_chain_axis_to_root, shard_bounds, shard_chains, world as world_fn
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert world_fn() == (rank, world)
        first, count = shard_chains(n_chains)
        assert (first, count) == shard_bounds(n_chains, rank, world)
        rows, entries = 3, 2
        g = torch.arange(first, first + count, dtype=torch.float64)
        local = torch.stack([torch.stack([g * 10 + r + 0.5 * e for e in range(entries)]) for r in range(rows)])    # [rows, entries, count]
        counts = [shard_bounds(n_chains, r, world)[1] for r in range(world)]
        full = all_gather_chain_axis(local, counts)
        gg = torch.arange(0, n_chains, dtype=torch.float64)
        want = torch.stack([torch.stack([gg * 10 + r + 0.5 * e for e in range(entries)]) for r in range(rows)])
        rooted = gather_chain_axis_to_root(local, counts, 0)
        ok_root = (rooted is None) if rank != 0 else bool(torch.equal(rooted, want))
        q.put((rank, bool(torch.equal(full, want)) and ok_root, tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [64, 37])          # equal blocks (all_gather_into_tensor) and ragged blocks (all_gather)
def test_sample_gather_over_gloo_world2(n_chains):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chains, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok, shape in res:
        assert ok and shape == (3, 2, n_chains), (rank, shape)
