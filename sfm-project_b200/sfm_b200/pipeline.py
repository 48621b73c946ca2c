"""The throughput path: match a pair list from the bank, then verify every pair (K2 -> K5 -> K4).

This is the batched counterpart of the reference's serial double loop (code/pipeline.py:36-49) plus the
geometric-verification stage the reference left empty (code/pipeline.py:60-65).  ``pipeline.py`` itself is
left untouched; this driver sits beside it.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .bank import DescriptorBank
from .matcher import MatchBatch, match_pairs
from .ransac import VerifyBatch, verify_corr


@dataclass
class VerifiedPairs:
    """Per-pair results, device resident until ``to_host``."""

    pairs: torch.Tensor          # int32 [P,2]
    n_matches: torch.Tensor      # int32 [P]
    matches: torch.Tensor        # int32 [P,cap,3]
    F: torch.Tensor              # float64 [P,3,3]
    n_inliers: torch.Tensor      # int32 [P]
    inlier_mask: torch.Tensor    # uint8 [P,cap]
    iters: torch.Tensor          # int32 [P]

    def to_host(self, with_matches: bool = True) -> dict:
        """One device->host transfer per array (the result a caller of the pipeline consumes)."""
        out = {
            "pairs": self.pairs.cpu().numpy(),
            "n_matches": self.n_matches.cpu().numpy(),
            "F": self.F.cpu().numpy(),
            "n_inliers": self.n_inliers.cpu().numpy(),
            "iters": self.iters.cpu().numpy(),
        }
        if with_matches:
            cap = self.matches.shape[1]
            live = torch.arange(cap, device=self.matches.device)[None, :] < self.n_matches[:, None]
            out["matches"] = self.matches[live].cpu().numpy()          # [sum n_matches, 3], pair-major
            out["inlier"] = self.inlier_mask[live].cpu().numpy()       # [sum n_matches]
            out["offsets"] = np.concatenate([[0], np.cumsum(out["n_matches"], dtype=np.int64)])
        return out


def match_and_verify(bank: DescriptorBank, pairs, *, ratio=0.75, ratio_mode="cv2_f32", mutual=False, impl="auto",
                     thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", score="sym_epipolar", lo=False, seed=0,
                     min_inliers=0, pair_batch: int = 2048, pair_ids=None) -> VerifiedPairs:
    """Match and verify every pair of ``pairs`` (int32 [P,2], image ids in the bank).
    ``pair_ids`` (default 0..P-1) name the RANSAC sample stream of each pair, so a sharded run that
    passes global pair indices reproduces the single-GPU result exactly."""
    pairs_host = np.asarray(pairs.cpu() if isinstance(pairs, torch.Tensor) else pairs, np.int32).reshape(-1, 2)
    P = pairs_host.shape[0]
    pair_ids = np.arange(P) if pair_ids is None else np.asarray(pair_ids).reshape(P)
    outs = []
    for s in range(0, max(P, 1), pair_batch):
        chunk = pairs_host[s: s + pair_batch]
        mb: MatchBatch = match_pairs(bank, chunk, ratio=ratio, ratio_mode=ratio_mode, mutual=mutual, impl=impl)
        vb: VerifyBatch = verify_corr(
            mb.corr, mb.counts, pair_id=pair_ids[s: s + len(chunk)], thr=thr, confidence=confidence,
            max_iters=max_iters, solver=solver, score=score, lo=lo, seed=seed, min_inliers=min_inliers)
        outs.append((mb, vb))
    cat = torch.cat
    return VerifiedPairs(
        pairs=cat([m.pairs for m, _ in outs]), n_matches=cat([m.counts for m, _ in outs]),
        matches=cat([m.matches for m, _ in outs]), F=cat([v.F for _, v in outs]),
        n_inliers=cat([v.n_inliers for _, v in outs]), inlier_mask=cat([v.mask for _, v in outs]),
        iters=cat([v.iters for _, v in outs]),
    )
