"""Posterior summaries formed on the device: pooled mean / sd, exact quantiles and the Gelman-Rubin
statistic per monitored entry, over all chains and kept rows of one `sample()` block that never leaves HBM.

The reference returns raw draws (mcmc.js:1029) and its README summarises them on the caller's side (README.md:44-52); at
2^20..2^22 chains the raw block is GBs per call, so `AmwgSampler.sample_summary(n)` keeps it on the GPU and moves a few
hundred bytes instead. Multi-GPU (one process per GPU): every rank reduces its own shard; the shards are combined with two
small collectives -- an all-gather of the per-rank moment records (merged exactly, in rank order) and a sum all-reduce of the
radix-select digit counts (integers) -- so every rank returns the same numbers as a single GPU holding all chains.

Host logic here is plain numpy (tested on CPU); the device work is behind `CudaBlockReducer` (C ABI: amwg_summary_moments,
amwg_summary_digit_hist). There is no CPU fallback: without the library or a GPU the reducer raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

MAX_PREFIXES = 32                      # include/amwg.h: n_prefix <= 32
_SIGN = np.uint64(1 << 63)


# ---------------------------------------------------------------------------------------------------------------------
# moments
def merge_moment_records(records: Sequence[np.ndarray]) -> np.ndarray:
    """Chan merge of per-shard records [entries, 4] = (chains, mean of chain means, M2 of chain means, sum of within-chain M2),
    in the order given (rank order): the same arithmetic as the device tree, so shards combine exactly."""
    acc = np.array(records[0], dtype=np.float64, copy=True)
    for rec in records[1:]:
        b = np.asarray(rec, dtype=np.float64)
        n = acc[:, 0] + b[:, 0]
        d = b[:, 1] - acc[:, 1]
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = np.where(b[:, 0] == 0, acc[:, 1], np.where(acc[:, 0] == 0, b[:, 1], acc[:, 1] + d * (b[:, 0] / n)))
            m2 = np.where(b[:, 0] == 0, acc[:, 2], np.where(acc[:, 0] == 0, b[:, 2], acc[:, 2] + b[:, 2] + d * d * (acc[:, 0] * b[:, 0] / n)))
        acc = np.stack([n, mean, m2, acc[:, 3] + b[:, 3]], axis=1)
    return acc


def finalize_moments(rec: np.ndarray, rows: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(mean, sd, rhat) per entry from the merged record. sd: pooled over all rows*chains draws, ddof=1. rhat: Gelman-Rubin
    potential scale reduction sqrt(((n-1)/n W + B/n) / W) with n = rows, W the mean within-chain variance, B/n the variance of
    the chain means (NaN with fewer than 2 rows or chains, or when W = 0)."""
    G, mean, b2, sw = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3]
    M = G * rows
    with np.errstate(invalid="ignore", divide="ignore"):
        sd = np.sqrt((sw + rows * b2) / (M - 1))
        W = sw / (G * (rows - 1))
        varplus = (rows - 1) / rows * W + b2 / (G - 1)
        rhat = np.sqrt(varplus / W)
    return mean, sd, rhat


# ---------------------------------------------------------------------------------------------------------------------
# exact quantiles: MSD radix select on the order-preserving key of an IEEE double
def key_to_double(keys: np.ndarray) -> np.ndarray:
    k = np.asarray(keys, dtype=np.uint64)
    u = np.where((k & _SIGN) != 0, k ^ _SIGN, ~k)
    return u.view(np.float64) if u.ndim else np.array([u], dtype=np.uint64).view(np.float64)[0]


def double_to_key(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
    return np.where((u & _SIGN) != 0, ~u, u | _SIGN)


def quantile_targets(M: int, probs: Sequence[float]):
    """numpy.quantile's default (linear) rule: position p*(M-1) between order statistics lo and lo+1.
    Returns (sorted distinct 0-based ranks, per-prob (index of lo, index of hi, fraction))."""
    ranks: List[int] = []
    plan = []
    for p in probs:
        if not (0.0 <= p <= 1.0):
            raise ValueError("probs must be in [0, 1]")
        pos = p * (M - 1)
        lo = int(np.floor(pos))
        hi = min(lo + 1, M - 1)
        plan.append((lo, hi, pos - lo))
        ranks += [lo, hi]
    uniq = sorted(set(ranks))
    index = {r: i for i, r in enumerate(uniq)}
    return np.asarray(uniq, dtype=np.int64), [(index[lo], index[hi], g) for lo, hi, g in plan]


def _lerp(a, b, t):
    """numpy's _lerp (lib/_function_base_impl.py): a + (b-a)*t, computed from b when t >= 0.5."""
    d = b - a
    return np.where(t >= 0.5, b - d * (1 - t), a + d * t)


class RadixSelect:
    """Host side of the 8-pass select: which prefixes the device should histogram next, and how the summed counts narrow
    each wanted order statistic down by one byte. ranks: sorted 0-based ranks, shared by all entries."""

    def __init__(self, entries: int, ranks: np.ndarray):
        self.E = entries
        self.T = len(ranks)
        self.prefix = np.zeros((entries, self.T), dtype=np.uint64)
        self.rem = np.tile(np.asarray(ranks, dtype=np.int64), (entries, 1))
        self.npass = 0

    def prefixes(self) -> Tuple[np.ndarray, np.ndarray]:
        """([entries, n_prefix] uint64 distinct prefixes padded with repeats, [entries, T] index of each target's prefix)."""
        uniq = [np.unique(self.prefix[e]) for e in range(self.E)]
        n = max(len(u) for u in uniq)
        if n > MAX_PREFIXES:
            raise ValueError("too many quantiles at once: %d distinct order statistics (max %d)" % (n, MAX_PREFIXES))
        table = np.empty((self.E, n), dtype=np.uint64)
        which = np.empty((self.E, self.T), dtype=np.int64)
        for e, u in enumerate(uniq):
            table[e, :len(u)] = u
            table[e, len(u):] = u[0]
            which[e] = np.searchsorted(u, self.prefix[e])
        return table, which

    def advance(self, counts: np.ndarray, which: np.ndarray) -> None:
        """counts [entries, n_prefix, 256] summed over all shards for the prefixes handed out by prefixes()."""
        sel = np.take_along_axis(np.asarray(counts, dtype=np.int64), which[:, :, None], axis=1)      # [E, T, 256]: each target's histogram
        cum = np.cumsum(sel, axis=2)
        d = (cum <= self.rem[:, :, None]).sum(axis=2)                                             # first byte whose cumulative count exceeds the rank
        if np.any(d > 255):
            raise RuntimeError("radix select: rank beyond the counted values (inconsistent histogram)")
        below = np.take_along_axis(cum, np.maximum(d - 1, 0)[:, :, None], axis=2)[:, :, 0]
        self.rem = self.rem - np.where(d > 0, below, 0)
        self.prefix = (self.prefix << np.uint64(8)) | d.astype(np.uint64)
        self.npass += 1

    def values(self) -> np.ndarray:
        assert self.npass == 8
        return key_to_double(self.prefix)


# ---------------------------------------------------------------------------------------------------------------------
class CudaBlockReducer:
    """The two device reductions over a torch CUDA tensor block[rows, entries, chains] (fp64, contiguous)."""

    def __init__(self, device: int):
        from . import _ffi
        self.L = _ffi.lib()                       # raises when the extension is missing: no CPU fallback
        self._ffi = _ffi
        self.device = device

    def moments(self, block) -> np.ndarray:
        rows, entries, chains = block.shape
        out = np.empty((entries, 4), dtype=np.float64)
        self._ffi.check(self.L.amwg_summary_moments(self.device, block.data_ptr(), rows, entries, chains, out.ctypes.data))
        return out

    def digit_counts(self, block, npass: int, prefix_table: np.ndarray):
        """-> torch int64 CUDA tensor [entries, n_prefix, 256] (this shard's counts)."""
        import torch
        rows, entries, chains = block.shape
        n_prefix = prefix_table.shape[1]
        dev = block.device
        pre = torch.from_numpy(prefix_table.view(np.int64).copy()).to(dev)
        counts = torch.zeros((entries, n_prefix, 256), dtype=torch.int64, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        self._ffi.check(self.L.amwg_summary_digit_hist(self.device, block.data_ptr(), rows, entries, chains, npass,
                                                       pre.data_ptr(), n_prefix, counts.data_ptr()))
        return counts


def summarise_block(reducer, block, rows: int, total_chains: int, probs: Sequence[float], distributed: bool):
    """-> (mean, sd, rhat, quantiles[len(probs)]) per entry, over all shards. `reducer` does the per-shard device work;
    the collectives run on the tensors it returns (NCCL for CUDA tensors, gloo for the CPU stand-in used in the tests)."""
    import torch
    entries = block.shape[1]
    rec = reducer.moments(block)
    if distributed:
        import torch.distributed as dist
        ws = dist.get_world_size()
        mine = torch.from_numpy(rec.copy())
        if block.is_cuda:
            mine = mine.to(block.device)
        gathered = torch.empty((ws * entries, 4), dtype=mine.dtype, device=mine.device)     # concatenated on dim 0
        dist.all_gather_into_tensor(gathered, mine)
        rec = merge_moment_records(list(gathered.cpu().numpy().reshape(ws, entries, 4)))
    mean, sd, rhat = finalize_moments(rec, rows)

    probs = [float(p) for p in probs]
    q = np.empty((len(probs), entries))
    per_select = MAX_PREFIXES // 2                            # every probability needs at most two order statistics
    for first in range(0, len(probs), per_select):            # long probability grids (equal-mass histograms): several selects
        chunk = probs[first:first + per_select]
        ranks, plan = quantile_targets(rows * total_chains, chunk)
        sel = RadixSelect(entries, ranks)
        for npass in range(8):
            table, which = sel.prefixes()
            counts = reducer.digit_counts(block, npass, table)
            if distributed:
                import torch.distributed as dist
                dist.all_reduce(counts)                       # integer sums: exact, independent of the number of GPUs
            sel.advance(counts.cpu().numpy(), which)
        vals = sel.values()                                   # [entries, T]
        for i, (lo, hi, g) in enumerate(plan):
            q[first + i] = _lerp(vals[:, lo], vals[:, hi], g)
    return mean, sd, rhat, q
