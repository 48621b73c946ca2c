"""Batched RANSAC fundamental-matrix verification (C ABI: sfm_ransac_f_batch).

Fills the reference's empty ``code/geometric_verification.py``; conventions follow
``cv2.findFundamentalMat(FM_RANSAC)``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class VerifyBatch:
    """F float64 [P,3,3] (zeros when no model), n_inliers int32 [P], mask uint8 [P, cap], iters int32 [P]."""

    F: torch.Tensor
    n_inliers: torch.Tensor
    mask: torch.Tensor
    iters: torch.Tensor


def ransac_params(*, thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", score="sym_epipolar", lo=False, seed=0,
                  min_inliers=0) -> _lib.RansacParams:
    if solver not in _lib.SOLVERS:
        raise ValueError(f"solver must be '7pt' or '8pt', got {solver!r}")
    if score not in _lib.SCORES:
        raise ValueError(f"score must be one of {list(_lib.SCORES)}, got {score!r}")
    if not (thr > 0) or max_iters <= 0:
        raise ValueError("thr and max_iters must be positive")
    p = _lib.RansacParams()
    p.solver, p.score = _lib.SOLVERS[solver], _lib.SCORES[score]
    p.threshold, p.max_iters, p.confidence = float(thr), int(max_iters), float(confidence)
    p.seed, p.lo_refit, p.min_inliers = int(seed), int(bool(lo)), int(min_inliers)
    return p


def verify_corr(corr: torch.Tensor, counts: torch.Tensor, *, pair_id=None, samples=None, **kw) -> VerifyBatch:
    """corr float32 [P, cap, 4] (x1,y1,x2,y2) on the GPU, counts int32 [P]."""
    if corr.dtype != torch.float32 or corr.dim() != 3 or corr.shape[2] != 4 or not corr.is_cuda:
        raise ValueError("corr must be a CUDA float32 tensor [P, cap, 4]")
    corr = corr.contiguous()
    counts = counts.to(device=corr.device, dtype=torch.int32).contiguous()
    P, cap, dev = corr.shape[0], corr.shape[1], corr.device
    prm = ransac_params(**kw)
    F = torch.zeros((P, 3, 3), dtype=torch.float64, device=dev)
    ninl = torch.zeros(P, dtype=torch.int32, device=dev)
    mask = torch.zeros((P, cap), dtype=torch.uint8, device=dev)
    iters = torch.zeros(P, dtype=torch.int32, device=dev)
    pid = None if pair_id is None else torch.as_tensor(np.asarray(pair_id, np.int64).astype(np.uint32).view(np.int32)).to(dev)
    smp = None
    if samples is not None:
        s = np.ascontiguousarray(samples, np.uint32)
        if s.shape != (prm.max_iters, 8):
            raise ValueError(f"samples must be [max_iters, 8], got {s.shape}")
        smp = torch.from_numpy(s.view(np.int32)).to(dev)
    if P:
        _lib.check(
            _lib.lib().sfm_ransac_f_batch(_lib.ptr(corr), cap, _lib.ptr(counts), P, _lib.ptr(pid), _lib.ptr(smp), C.byref(prm),
                                          _lib.ptr(F), _lib.ptr(ninl), _lib.ptr(mask), _lib.ptr(iters),
                                          _lib.current_stream_ptr(dev)),
            "sfm_ransac_f_batch",
        )
    return VerifyBatch(F, ninl, mask, iters)
