"""The throughput path: match a pair list from the bank, then verify every pair (K2 -> K5 -> K4).

This is the batched counterpart of the reference's serial double loop (code/pipeline.py:36-49) plus the
geometric-verification stage the reference left empty (code/pipeline.py:60-65).  ``pipeline.py`` itself is
left untouched; this driver sits beside it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from .bank import DescriptorBank
from .plan import HotPathPlan


@dataclass
class VerifiedPairs:
    """Per-pair summaries stay on the device; the variable-length part (matches, inlier flags) is streamed to the
    host batch by batch when ``fetch=True`` and is then available through ``to_host``."""

    pairs: torch.Tensor          # int32 [P,2]
    n_matches: torch.Tensor      # int32 [P]
    F: torch.Tensor              # float64 [P,3,3]  (zeros when no model)
    n_inliers: torch.Tensor      # int32 [P]
    iters: torch.Tensor          # int32 [P]        hypotheses evaluated
    host: dict | None = field(default=None, repr=False)
    d2h_bytes: int = 0

    def to_host(self, with_matches: bool = True) -> dict:
        """``pairs, n_matches, F, n_inliers, iters`` and, with ``with_matches``, the packed rows
        ``matches`` int32 [sum n_matches, 3] = (queryIdx, trainIdx, squared L2), ``inlier`` uint8 [sum n_matches],
        ``offsets`` int64 [P+1] (pair p owns rows offsets[p]:offsets[p+1])."""
        if self.host is not None:
            out = dict(self.host)
            if not with_matches:
                for k in ("matches", "inlier", "offsets"):
                    out.pop(k, None)
            return out
        if with_matches:
            raise ValueError("matches were not fetched: call match_and_verify(..., fetch=True)")
        return {"pairs": self.pairs.cpu().numpy(), "n_matches": self.n_matches.cpu().numpy(), "F": self.F.cpu().numpy(),
                "n_inliers": self.n_inliers.cpu().numpy(), "iters": self.iters.cpu().numpy()}


_PLAN_KEYS = ("ratio", "ratio_mode", "mutual", "impl", "thr", "confidence", "max_iters", "solver", "score", "lo", "seed",
              "min_inliers", "prefilter")


def get_plan(bank: DescriptorBank, batch: int, **params) -> HotPathPlan:
    """Plans (device + pinned buffers) are cached on the bank, one per (batch size, parameter set)."""
    key = (int(batch),) + tuple(params.get(k) for k in _PLAN_KEYS)
    cache = bank.__dict__.setdefault("_plans", {})
    plan = cache.get(key)
    if plan is None:
        if len(cache) >= 4:
            cache.clear()
        plan = cache[key] = HotPathPlan(bank, batch, **params)
    return plan


def match_and_verify(bank: DescriptorBank, pairs, *, ratio=0.75, ratio_mode="cv2_f32", mutual=False, impl="auto",
                     thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", score="sym_epipolar", lo=False, seed=0,
                     min_inliers=0, pair_batch: int = 2048, pair_ids=None, fetch=False,
                     prefilter: bool = True) -> VerifiedPairs:
    """Match and verify every pair of ``pairs`` (int32 [P,2], image ids in the bank).

    ``pair_ids`` (default 0..P-1) name the RANSAC sample stream of each pair, so a sharded run that passes global
    pair indices reproduces the single-GPU result exactly.  ``fetch=True`` also streams every batch's matches and
    inlier flags to the host (pinned buffers, copies overlapped with the RANSAC kernel of the same batch and the
    sweep of the next one); ``fetch="view"`` hands out the pinned buffers themselves when the run is a single batch
    (zero host copies; valid until the next call on this bank).  ``prefilter`` lets the sweep drop rows that provably fail the ratio test before the
    exact refinement (results are identical with or without it)."""
    pairs_host = np.ascontiguousarray(np.asarray(pairs.cpu() if isinstance(pairs, torch.Tensor) else pairs, np.int32).reshape(-1, 2))
    P = pairs_host.shape[0]
    if P and (pairs_host.min() < 0 or pairs_host.max() >= bank.n_images):
        raise ValueError(f"pair list refers to images outside [0, {bank.n_images})")
    ids_host = np.arange(P, dtype=np.int64) if pair_ids is None else np.asarray(pair_ids, np.int64).reshape(P)
    dev = bank.device
    if P == 0:
        z = lambda *s, dt=torch.int32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
        host = None
        if fetch:
            host = {"pairs": pairs_host, "n_matches": np.zeros(0, np.int32), "F": np.zeros((0, 3, 3)), "n_inliers": np.zeros(0, np.int32),
                    "iters": np.zeros(0, np.int32), "matches": np.zeros((0, 3), np.int32), "inlier": np.zeros(0, np.uint8),
                    "offsets": np.zeros(1, np.int64)}
        return VerifiedPairs(z(0, 2), z(0), z(0, 3, 3, dt=torch.float64), z(0), z(0), host)
    batch = int(min(pair_batch, P))
    plan = get_plan(bank, batch, ratio=ratio, ratio_mode=ratio_mode, mutual=mutual, impl=impl, thr=thr, confidence=confidence,
                    max_iters=max_iters, solver=solver, score=score, lo=lo, seed=seed, min_inliers=min_inliers, prefilter=prefilter)
    # one upload of the whole pair list and its RANSAC stream ids (pinned -> device, asynchronous)
    pairs_d = torch.from_numpy(pairs_host).pin_memory().to(dev, non_blocking=True)
    ids_d = torch.from_numpy(ids_host.astype(np.uint32).view(np.int32)).pin_memory().to(dev, non_blocking=True)
    rev_d = pairs_d.flip(1).contiguous() if mutual else None
    n_matches = torch.empty(P, dtype=torch.int32, device=dev)
    F = torch.empty((P, 3, 3), dtype=torch.float64, device=dev)
    n_inl = torch.empty(P, dtype=torch.int32, device=dev)
    iters = torch.empty(P, dtype=torch.int32, device=dev)
    chunks, d2h = [], 0

    def collect(s, n):
        nonlocal d2h
        h = plan.fetch_end()
        d2h += plan.host_bytes()
        # the pinned buffers are reused by the next batch / call: copy out unless the caller asked for views
        chunks.append(h if (fetch == "view" and P <= batch) else {k: v.copy() for k, v in h.items()})

    pending = None
    for s in range(0, P, batch):
        n = min(batch, P - s)
        plan.launch(pairs_d[s: s + n], ids_d[s: s + n], None if rev_d is None else rev_d[s: s + n])
        # per-pair summaries of this batch into the full-length device arrays (tiny device copies, stream ordered)
        n_matches[s: s + n].copy_(plan.counts[:n])
        F[s: s + n].copy_(plan.F[:n])
        n_inl[s: s + n].copy_(plan.ninl[:n])
        iters[s: s + n].copy_(plan.iters[:n])
        if fetch:
            if pending is not None:
                collect(*pending)              # previous batch's copies finished long ago (they overlapped this sweep)
            plan.fetch_begin()
            pending = (s, n)
    host = None
    if fetch:
        collect(*pending)
        off = np.zeros(P + 1, np.int64)
        np.cumsum(np.concatenate([c["n_matches"] for c in chunks]), out=off[1:])
        cat = (lambda k: chunks[0][k]) if len(chunks) == 1 else (lambda k: np.concatenate([c[k] for c in chunks]))
        host = {"pairs": pairs_host, "n_matches": cat("n_matches"), "F": cat("F"), "n_inliers": cat("n_inliers"),
                "iters": cat("iters"), "matches": cat("matches"), "inlier": cat("inlier"), "offsets": off}
    return VerifiedPairs(pairs_d, n_matches, F, n_inl, iters, host, d2h)
