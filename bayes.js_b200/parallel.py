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
    """Distributed sample(): this rank's shard is sampled in row chunks; while chunk k+1 is being computed, chunk k is collected
    over the chain axis (NCCL over NVLink) and copied into a pinned host buffer on a side stream. `sampler.gather`: "all"
    (default; every rank returns all chains), "root" (rank 0 returns all chains, the others their own shard), "none" (own shard)."""
    import torch
    import torch.distributed as dist
    from . import _ffi
    from .mcmc import _pinned_empty
    from .tracer import JsThrow
    if not (dist.is_available() and dist.is_initialized()):
        raise JsThrow("options.distributed needs an initialised torch.distributed process group")
    L = _ffi.lib()
    dev = torch.device("cuda", sampler.device)
    ws, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_bounds(sampler.n_chains, r, ws)[1] for r in range(ws)]
    mode = getattr(sampler, "gather", "all")
    collect = mode == "all" or mode == "root"
    out_chains = sampler.n_chains if (mode == "all" or (mode == "root" and rank == 0)) else sampler.local_chains
    host = _pinned_empty((rows, len(mon), out_chains))
    host_t = torch.from_numpy(host)
    monp = mon.ctypes.data_as(C.POINTER(C.c_int32))
    chunk_rows = max(1, min(rows, 10))
    side = torch.cuda.Stream(device=dev)
    done_events = []
    row0 = 0
    while row0 < rows:
        r = min(chunk_rows, rows - row0)
        n_chunk = min(n - row0 * thin, r * thin)               # sample(a) then sample(b) == sample(a+b) when a is a multiple of thin
        local = torch.empty((r, len(mon), sampler.local_chains), dtype=torch.float64, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        rc = L.amwg_sample_device(sampler._handle, n_chunk, thin, monp, len(mon), local.data_ptr())      # blocks until the chunk is in HBM
        if rc != 0:
            raise JsThrow(L.amwg_last_error().decode())
        with torch.cuda.stream(side):                            # collect + D2H of this chunk overlap the next chunk's sweeps
            if mode == "all":
                full = all_gather_chain_axis(local, counts)
            elif mode == "root":
                full = gather_chain_axis_to_root(local, counts, 0)
                if full is None:
                    full = local
            else:
                full = local
            host_t[row0:row0 + r].copy_(full, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
            done_events.append((ev, local, full))                # keep the device buffers alive until the copy has finished
        row0 += r
    side.synchronize()
    del done_events, collect
    return host
