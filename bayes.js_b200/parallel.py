"""Multi-GPU plumbing: one process per GPU (torch.distributed), chains sharded, no data-path collective.

Chains are independent (the reference runs exactly one, mcmc.js:940-966), so the path shards embarrassingly: rank r
owns the contiguous block of global chain ids [r*C/G, (r+1)*C/G).  The Philox stream is keyed by the GLOBAL chain id,
so the draws do not depend on G.  The only collective is the final all-gather of the sample blocks in `sample()`
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import numpy as np


def world() -> Tuple[int, int]:
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun environment."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_bounds(n_chains: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: (first_chain, count). Remainder chains go to the lowest ranks."""
    base, rem = divmod(int(n_chains), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_chains(n_chains: int) -> Tuple[int, int]:
    rank, ws = world()
    return shard_bounds(n_chains, rank, ws)


def all_gather_chain_axis(local, counts):
    """All-gather `local[rows, entries, c_rank]` over its last (chain) axis -> `[rows, entries, sum(counts)]`, chains in
    global order. One collective per (row, entry) slab, written straight into its final place (the kernel's layout keeps the
    chain axis fastest, so a single dim-0 all-gather would interleave ranks). CUDA tensors over NCCL, CPU tensors over gloo."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size()
    rows, entries = local.shape[0], local.shape[1]
    cmax = max(counts)
    ragged = len(set(counts)) > 1
    if ragged and local.shape[-1] < cmax:              # pad to the largest block (collectives want equal sizes)
        pad = torch.zeros((rows, entries, cmax - local.shape[-1]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=-1)
    local = local.contiguous()
    out = torch.empty((rows, entries, ws * cmax), dtype=local.dtype, device=local.device)
    for r in range(rows):
        for e in range(entries):
            dist.all_gather_into_tensor(out[r, e], local[r, e])
    if not ragged:
        return out
    keep = torch.cat([torch.arange(k * cmax, k * cmax + c, device=local.device) for k, c in enumerate(counts)])
    return out.index_select(-1, keep)


def gather_chain_axis_to_root(local, counts, root: int = 0):
    """Gather `local[rows, entries, c_rank]` over the chain axis onto rank `root` only (returns None elsewhere):
    the host-memory-friendly variant when one process collects the draws."""
    import torch
    import torch.distributed as dist
    ws, rank = dist.get_world_size(), dist.get_rank()
    rows, entries = local.shape[0], local.shape[1]
    cmax = max(counts)
    if len(set(counts)) > 1:
        full = all_gather_chain_axis(local, counts)          # ragged shards: reuse the padded all-gather
        return full if rank == root else None
    local = local.contiguous()
    out = torch.empty((rows, entries, ws * cmax), dtype=local.dtype, device=local.device) if rank == root else None
    for r in range(rows):
        for e in range(entries):
            dist.gather(local[r, e], list(out[r, e].view(ws, cmax).unbind(0)) if rank == root else None, dst=root)
    return out


def sample_and_gather(sampler, n: int, thin: int, mon: np.ndarray, rows: int) -> np.ndarray:
    """sample() on this rank's shard into a device buffer, collect over the chain axis (NCCL over NVLink), then one D2H into
    a pinned host buffer. `sampler.gather`: "all" (default; every rank returns all chains), "root" (rank 0 returns all chains,
    the other ranks their own shard), "none" (every rank returns its own shard)."""
    import torch
    import torch.distributed as dist
    from . import _ffi
    from .mcmc import _pinned_empty
    from .tracer import JsThrow
    if not (dist.is_available() and dist.is_initialized()):
        raise JsThrow("options.distributed needs an initialised torch.distributed process group")
    L = _ffi.lib()
    dev = torch.device("cuda", sampler.device)
    local = torch.empty((rows, len(mon), sampler.local_chains), dtype=torch.float64, device=dev)
    torch.cuda.synchronize(dev)
    rc = L.amwg_sample_device(sampler._handle, n, thin, mon.ctypes.data_as(C.POINTER(C.c_int32)), len(mon), local.data_ptr())
    if rc != 0:
        raise JsThrow(L.amwg_last_error().decode())
    ws = dist.get_world_size()
    counts = [shard_bounds(sampler.n_chains, r, ws)[1] for r in range(ws)]
    mode = getattr(sampler, "gather", "all")
    if mode == "all":
        full = all_gather_chain_axis(local, counts)
    elif mode == "root":
        full = gather_chain_axis_to_root(local, counts, 0)
        if full is None:
            full = local
    else:
        full = local
    host = _pinned_empty(tuple(full.shape))
    torch.from_numpy(host).copy_(full, non_blocking=True)
    torch.cuda.synchronize(dev)
    return host
