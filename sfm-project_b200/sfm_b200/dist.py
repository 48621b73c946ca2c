"""Pair-sharded multi-GPU execution (one process per GPU, torch.distributed).

The reference's pair loop (code/pipeline.py:38-47) is embarrassingly parallel by image pair, so the
only communication is (1) one broadcast of the packed descriptor bank from rank 0 and (2) one gather
of per-pair results to rank 0.  No collective runs inside the compute phase, so there
is nothing to fuse with a kernel.  Everything here is device-agnostic plumbing: the same code runs
under NCCL on GPUs and under gloo on CPU tensors (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition(n_items: int, rank: int, world_size: int, mode: str = "block"):
    """Index array of the items rank ``rank`` owns.

    block : contiguous ceil(P/R) slices of the (i,j)-sorted pair list (keeps image i hot in L2)
    cyclic: items rank, rank+R, ... (balances RANSAC cost when match counts vary along the list)
    """
    if mode == "block":
        per = -(-n_items // world_size)
        return np.arange(min(rank * per, n_items), min((rank + 1) * per, n_items))
    if mode == "cyclic":
        return np.arange(rank, n_items, world_size)
    raise ValueError(f"unknown partition mode {mode!r}")


def broadcast_bank(bank, src: int = 0):
    """Broadcast the packed bank storage (descriptors, K-extension, norms, xy, counts) from ``src``.
    Non-source ranks end up with a ready bank without running the pack kernel."""
    rank, ws = world()
    meta = torch.tensor([bank.n_images if rank == src else 0], dtype=torch.int64, device=bank.storage.device)
    if ws > 1:
        dist.broadcast(bank.storage, src=src)
        dist.broadcast(meta, src=src)
    if rank != src:
        n = int(meta.item())
        bank.mark_filled(n, bank.counts[:n].cpu().numpy())
    return bank


def gather_varlen(t: torch.Tensor, dst: int = 0):
    """Gather tensors whose first dimension differs per rank.  Returns the concatenation (rank order) on
    ``dst`` and None elsewhere.  Sizes travel first (all_gather), payloads are padded to the maximum."""
    rank, ws = world()
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    if dist.get_backend() == "nccl":
        # NCCL gather is implemented with grouped send/recv; all_gather of padded buffers is the simple exact form
        bufs = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(bufs, pad)
    else:
        bufs = [torch.empty_like(pad) for _ in range(ws)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def gather_pair_results(local: dict, order: np.ndarray, n_total: int, dst: int = 0):
    """``local`` maps name -> tensor with one row per locally owned pair; ``order`` are the global pair
    indices of those rows.  Rank ``dst`` receives every array re-assembled in global pair order."""
    rank, ws = world()
    dev = next(iter(local.values())).device
    idx = gather_varlen(torch.as_tensor(order, dtype=torch.int64, device=dev), dst)
    out = {}
    for name, t in local.items():
        g = gather_varlen(t, dst)
        if rank == dst:
            full = torch.zeros((n_total,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
            full[idx] = g
            out[name] = full
    return out if rank == dst else None


def gather_summaries(res, order: np.ndarray, n_total: int, dst: int = 0):
    """Per-pair summaries (n_matches, n_inliers, iters, F -- plus H, n_inliers_h and R, t, n_pose when the optional stages
    ran) of every rank on ``dst`` in global pair order with ONE collective and no host synchronisation: the block/cyclic
    partition sizes are known on every rank, so each rank contributes a fixed-size float64 [ceil(P/R), W] tile (integers
    are exact in float64).  Returns a dict on ``dst``."""
    rank, ws = world()
    dev = res.n_matches.device
    n = len(order)
    per = -(-n_total // ws)
    # column layout: name -> (first column, width, dtype, trailing shape)
    cols, w = {}, 1
    def add(name, width, dt, shape):
        nonlocal w
        cols[name] = (w, width, dt, shape)
        w += width
    add("n_matches", 1, torch.int32, ())
    add("n_inliers", 1, torch.int32, ())
    add("iters", 1, torch.int32, ())
    add("F", 9, torch.float64, (3, 3))
    if getattr(res, "H", None) is not None:
        add("H", 9, torch.float64, (3, 3))
        add("n_inliers_h", 1, torch.int32, ())
    if getattr(res, "R", None) is not None:
        add("R", 9, torch.float64, (3, 3))
        add("t", 3, torch.float64, (3,))
        add("n_pose", 1, torch.int32, ())
    tile = torch.zeros((per, w), dtype=torch.float64, device=dev)
    tile[:n, 0] = torch.as_tensor(order, dtype=torch.float64, device=dev) + 1.0          # 0 marks padding
    for name, (c0, width, _, _) in cols.items():
        tile[:n, c0: c0 + width] = getattr(res, name).to(torch.float64).reshape(n, width)
    if ws == 1:
        full = tile
    else:
        full = torch.empty((ws * per, w), dtype=torch.float64, device=dev)
        try:
            dist.all_gather_into_tensor(full, tile)
        except (RuntimeError, NotImplementedError):                  # backends without the flat form
            parts = [torch.empty_like(tile) for _ in range(ws)]
            dist.all_gather(parts, tile)
            full = torch.cat(parts)
    if rank != dst:
        return None
    live = full[:, 0] > 0
    idx = (full[live, 0] - 1.0).to(torch.int64)
    out = {}
    for name, (c0, width, dt, shape) in cols.items():
        o = torch.zeros((n_total,) + shape, dtype=dt, device=dev)
        o[idx] = full[live, c0: c0 + width].to(dt).reshape((-1,) + shape)
        out[name] = o
    return out


def match_and_verify_sharded(bank, pairs, *, mode: str = "block", dst: int = 0, **params):
    """Every rank holds the (broadcast) bank; the pair list is partitioned; per-pair summaries
    (n_matches, n_inliers, F, iters; H / R / t / counts of the optional stages when requested) are gathered on ``dst`` in
    global pair order.
    Returns (gathered dict or None, local VerifiedPairs)."""
    from .pipeline import match_and_verify

    rank, ws = world()
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    mine = partition(len(pairs), rank, ws, mode)
    # pair_id = global pair index, so the RANSAC sample streams (and hence the results) do not depend on R
    res = match_and_verify(bank, pairs[mine], pair_ids=mine, **params)
    return gather_summaries(res, mine, len(pairs), dst), res
