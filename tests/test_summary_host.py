"""Host side of the on-device posterior summaries: exact radix-select bookkeeping, moment merging and the
collectives, on CPU tensors with a numpy stand-in for the two device reductions (tests/summary_ref.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBS = (0.0, 0.025, 0.25, 0.5, 0.75, 0.975, 1.0)


def _block(rows, entries, chains, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, (rows, entries, chains))
    x[:, 0] = 184.5 + 0.14 * x[:, 0]                       # a config-2-like mu: narrow and far from 0
    if entries > 1:
        x[:, 1] = np.round(x[:, 1] * 3)                     # an integer parameter: many ties, both signs, -0/+0
    if entries > 2:
        x[:, 2] = np.exp(5 * x[:, 2]) * np.sign(rng.normal(size=(rows, chains)))     # 20 orders of magnitude, both signs
    return x


def test_ordered_key_is_monotone_and_invertible(pkg):
    from bayes_js_b200.summary import double_to_key, key_to_double
    x = np.array([-np.inf, -1e300, -2.5, -1e-310, -0.0, 0.0, 5e-324, 1.0, 184.5, 1e300, np.inf])
    k = double_to_key(x)
    assert np.all(k[1:] > k[:-1])
    assert np.array_equal(key_to_double(k).view(np.uint64), x.view(np.uint64))


@pytest.mark.parametrize("rows,entries,chains", [(7, 3, 41), (1, 2, 64), (40, 1, 3), (2, 3, 1)])
def test_summary_of_one_shard_matches_numpy(pkg, rows, entries, chains):
    import torch
    from bayes_js_b200.summary import summarise_block
    from summary_ref import NumpyBlockReducer, numpy_summary
    x = _block(rows, entries, chains, rows * 100 + chains)
    mean, sd, rhat, q = summarise_block(NumpyBlockReducer(), torch.from_numpy(x), rows, chains, PROBS, False)
    m0, s0, r0, q0 = numpy_summary(x, PROBS)
    assert np.allclose(mean, m0, rtol=1e-13, atol=0)
    if rows * chains > 1:
        assert np.allclose(sd, s0, rtol=1e-11, atol=0)
    assert np.allclose(rhat, r0, rtol=1e-9, atol=0, equal_nan=True)
    assert np.array_equal(q, q0)                            # exact order statistics, numpy's interpolation rule


def test_long_probability_grids_run_as_several_selects(pkg):
    """more than 16 probabilities (an equal-mass histogram): split into selects of at most 32 order statistics each"""
    import torch
    from bayes_js_b200.summary import summarise_block
    from summary_ref import NumpyBlockReducer, numpy_summary
    x = _block(11, 2, 97, 8)
    probs = tuple(np.linspace(0, 1, 41))
    _, _, _, q = summarise_block(NumpyBlockReducer(), torch.from_numpy(x), 11, 97, probs, False)
    assert q.shape == (41, 2) and np.array_equal(q, numpy_summary(x, probs)[3])


def test_shards_merge_exactly(pkg):
    import torch
    from bayes_js_b200.summary import RadixSelect, finalize_moments, merge_moment_records, quantile_targets
    from summary_ref import NumpyBlockReducer, numpy_summary
    rows, entries, chains = 9, 3, 50
    x = _block(rows, entries, chains, 5)
    cuts = [0, 17, 18, 50]                                  # ragged shards, one of a single chain
    red = NumpyBlockReducer()
    shards = [torch.from_numpy(np.ascontiguousarray(x[:, :, a:b])) for a, b in zip(cuts, cuts[1:])]
    rec = merge_moment_records([red.moments(s) for s in shards])
    mean, sd, rhat = finalize_moments(rec, rows)
    m0, s0, r0, q0 = numpy_summary(x, PROBS)
    assert np.allclose(mean, m0, rtol=1e-13) and np.allclose(sd, s0, rtol=1e-11) and np.allclose(rhat, r0, rtol=1e-9)
    ranks, plan = quantile_targets(rows * chains, PROBS)
    sel = RadixSelect(entries, ranks)
    for p in range(8):
        table, which = sel.prefixes()
        sel.advance(sum(red.digit_counts(s, p, table).numpy() for s in shards), which)
    vals = sel.values()
    flat = np.sort(np.moveaxis(x, 1, 0).reshape(entries, -1), axis=1)
    assert np.array_equal(vals.view(np.uint64), flat[:, ranks].view(np.uint64))


def test_quantile_targets_and_limits(pkg):
    from bayes_js_b200.summary import MAX_PREFIXES, RadixSelect, quantile_targets
    ranks, plan = quantile_targets(11, [0.0, 0.5, 1.0, 0.33])
    assert list(ranks) == [0, 1, 3, 4, 5, 6, 10]
    assert plan[1] == (4, 5, 0.0) and plan[2][0] == 6 and plan[2][1] == 6
    with pytest.raises(ValueError):
        quantile_targets(10, [1.5])
    sel = RadixSelect(1, np.arange(MAX_PREFIXES + 1))
    sel.prefix[0] = np.arange(MAX_PREFIXES + 1, dtype=np.uint64)
    with pytest.raises(ValueError):
        sel.prefixes()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    graft.load_package()
    from bayes_js_b200.parallel import shard_bounds
    from bayes_js_b200.summary import summarise_block
    from summary_ref import NumpyBlockReducer, numpy_summary
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, entries, chains = 6, 3, 37                       # ragged: 19 + 18 chains
        x = _block(rows, entries, chains, 11)
        first, count = shard_bounds(chains, rank, world)
        mine = torch.from_numpy(np.ascontiguousarray(x[:, :, first:first + count]))
        mean, sd, rhat, qq = summarise_block(NumpyBlockReducer(), mine, rows, chains, PROBS, True)
        m0, s0, r0, q0 = numpy_summary(x, PROBS)
        ok = (np.allclose(mean, m0, rtol=1e-13) and np.allclose(sd, s0, rtol=1e-11) and np.allclose(rhat, r0, rtol=1e-9)
              and np.array_equal(qq, q0))
        q.put((rank, bool(ok), mean.tobytes() + sd.tobytes() + qq.tobytes()))
    finally:
        dist.destroy_process_group()


def test_summary_over_gloo_world2():
    """the N>1 path: every rank reduces its shard, the records are all-gathered and the digit counts all-reduced; both ranks end
    with the single-process numbers, bit for bit the same on both."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]
