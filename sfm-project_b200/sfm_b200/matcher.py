"""Batched pair matching on the bank (C ABI: sfm_match_knn2, sfm_filter_matches, sfm_match_hamming).

``match_pairs`` replaces the per-pair body of the reference loop (code/pipeline.py:38-47 calling
``extract_and_match``, code/feature_matching.py:41-60) for a whole pair list at once.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from fractions import Fraction

import numpy as np
import torch

from . import _lib
from .bank import DescriptorBank


@dataclass
class MatchBatch:
    """Device-resident result of a batched match.

    counts  int32 [P]            surviving matches per pair
    matches int32 [P, cap, 3]    (queryIdx, trainIdx, distance) -- squared L2 / Hamming; rows >= count undefined
    corr    float32 [P, cap, 4]  (x1, y1, x2, y2) or None
    knn     int32 [P, cap, 4]    (idx1, D1, idx2, D2) per query row, only when requested
    """

    pairs: torch.Tensor
    counts: torch.Tensor
    matches: torch.Tensor
    corr: torch.Tensor | None = None
    knn: torch.Tensor | None = None

    def to_host(self):
        """List of (q, t, d) int32 numpy triples, one per pair."""
        counts = self.counts.cpu().numpy()
        m = self.matches.cpu().numpy()
        return [(m[p, : counts[p], 0].copy(), m[p, : counts[p], 1].copy(), m[p, : counts[p], 2].copy()) for p in range(len(counts))]


def _pairs_tensor(pairs, device) -> torch.Tensor:
    if isinstance(pairs, torch.Tensor):
        return pairs.to(device=device, dtype=torch.int32).reshape(-1, 2).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))).to(device)


def _check_pairs(pairs_host: np.ndarray, bank: DescriptorBank) -> None:
    if pairs_host.size and (pairs_host.min() < 0 or pairs_host.max() >= bank.n_images):
        raise ValueError(f"pair list refers to images outside [0, {bank.n_images})")


def knn2(bank: DescriptorBank, pairs, impl: str = "auto", out: torch.Tensor | None = None, grid: int = 0,
         sweep_only: bool = False) -> torch.Tensor:
    """kNN(k=2) of every query row: int32 [P, feat_stride, 4] = (idx1, D1, idx2, D2)."""
    pairs_t = _pairs_tensor(pairs, bank.device)
    P = pairs_t.shape[0]
    if out is None:
        out = torch.empty((P, bank.feat_stride, 4), dtype=torch.int32, device=bank.device)
    prm = _lib.MatchParams()
    prm.impl, prm.grid = _lib.MATCH_IMPLS[impl], int(grid)
    prm.sweep_only = int(sweep_only)              # diagnostics: leave candidate records, skip the refinement (5/6: bounds)
    _lib.check(
        _lib.lib().sfm_match_knn2(bank.handle, _lib.ptr(pairs_t), P, C.byref(prm), _lib.ptr(out),
                                  _lib.current_stream_ptr(bank.device)),
        "sfm_match_knn2",
    )
    return out


def filter_params(ratio, ratio_mode, mutual, max_distance_sq=0) -> _lib.FilterParams:
    prm = _lib.FilterParams()
    if ratio is None:
        ratio_mode = None
    if ratio_mode not in _lib.RATIO_MODES:
        raise ValueError(f"ratio_mode must be one of {list(_lib.RATIO_MODES)}, got {ratio_mode!r}")
    prm.ratio_mode = _lib.RATIO_MODES[ratio_mode]
    prm.mutual = int(bool(mutual))
    prm.ratio = float(ratio) if ratio is not None else 1.0
    f = Fraction(prm.ratio).limit_denominator(1024)
    prm.ratio_num, prm.ratio_den = f.numerator, f.denominator
    prm.max_distance_sq = int(max_distance_sq)
    return prm


def match_pairs(bank: DescriptorBank, pairs, *, ratio=0.75, ratio_mode="cv2_f32", mutual=False, impl="auto",
                return_knn=False, with_corr=True, max_distance_sq=0) -> MatchBatch:
    """L2 kNN(k=2) + Lowe ratio (+ mutual nearest) for every pair; matches in ascending queryIdx."""
    if bank.metric != "l2":
        raise ValueError("match_pairs needs an L2 bank; use match_pairs_hamming for binary descriptors")
    pairs_host = np.asarray(pairs.cpu() if isinstance(pairs, torch.Tensor) else pairs, np.int32).reshape(-1, 2)
    _check_pairs(pairs_host, bank)
    pairs_t = _pairs_tensor(pairs, bank.device)
    P, cap, dev = pairs_t.shape[0], bank.feat_stride, bank.device
    counts = torch.zeros(P, dtype=torch.int32, device=dev)
    matches = torch.empty((P, cap, 3), dtype=torch.int32, device=dev)
    corr = torch.empty((P, cap, 4), dtype=torch.float32, device=dev) if with_corr else None
    if P == 0:
        return MatchBatch(pairs_t, counts, matches, corr, None)
    fwd = knn2(bank, pairs_t, impl)
    rev = knn2(bank, pairs_t.flip(1).contiguous(), impl) if mutual else None
    prm = filter_params(ratio, ratio_mode, mutual, max_distance_sq)
    _lib.check(
        _lib.lib().sfm_filter_matches(bank.handle, _lib.ptr(pairs_t), P, _lib.ptr(fwd), _lib.ptr(rev), C.byref(prm),
                                      _lib.ptr(counts), _lib.ptr(matches), _lib.ptr(corr), _lib.current_stream_ptr(dev)),
        "sfm_filter_matches",
    )
    return MatchBatch(pairs_t, counts, matches, corr, fwd if return_knn else None)


def match_pairs_packed(bank: DescriptorBank, pairs, *, ratio=0.75, ratio_mode="cv2_f32", mutual=False, prefilter=True, fused=True,
                       max_distance_sq=0):
    """The packed matcher of the throughput path on its own: (counts int32 [P], offsets int32 [P+1], matches int32 [total,3],
    corr float32 [total,4]).  ``fused`` = sfm_match_pairs_packed (refinement and filter in one pass); otherwise the kNN table
    path (sfm_match_knn2 + sfm_filter_matches_packed).  Both return the same arrays."""
    pairs_host = np.asarray(pairs.cpu() if isinstance(pairs, torch.Tensor) else pairs, np.int32).reshape(-1, 2)
    _check_pairs(pairs_host, bank)
    pairs_t = _pairs_tensor(pairs, bank.device)
    P, cap, dev = pairs_t.shape[0], bank.feat_stride, bank.device
    L, st = _lib.lib(), _lib.current_stream_ptr(dev)
    fprm = filter_params(ratio, ratio_mode, mutual, max_distance_sq)
    mprm = _lib.MatchParams()
    if prefilter and fprm.ratio_mode != _lib.RATIO_NONE:
        mprm.prefilter_mode, mprm.prefilter_ratio = fprm.ratio_mode, fprm.ratio
        mprm.prefilter_num, mprm.prefilter_den = int(fprm.ratio_num), int(fprm.ratio_den)
    counts = torch.zeros(P, dtype=torch.int32, device=dev)
    offsets = torch.zeros(P + 1, dtype=torch.int32, device=dev)
    matches = torch.empty((max(P * cap, 1), 3), dtype=torch.int32, device=dev)
    corr = torch.empty((max(P * cap, 1), 4), dtype=torch.float32, device=dev)
    knn = torch.empty((max(P, 1), cap, 4), dtype=torch.int32, device=dev)
    rev = knn2(bank, pairs_t.flip(1).contiguous()) if (mutual and P) else None
    if P == 0:
        return counts, offsets, matches[:0], corr[:0]
    if fused:
        blk = torch.empty(max(P, 1) * (cap // 256), dtype=torch.int32, device=dev)
        _lib.check(L.sfm_match_pairs_packed(bank.handle, _lib.ptr(pairs_t), P, C.byref(mprm), C.byref(fprm), _lib.ptr(rev), _lib.ptr(knn),
                                            _lib.ptr(blk), _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(matches), _lib.ptr(corr), st),
                   "sfm_match_pairs_packed")
    else:
        if P:
            _lib.check(L.sfm_match_knn2(bank.handle, _lib.ptr(pairs_t), P, C.byref(mprm), _lib.ptr(knn), st), "sfm_match_knn2")
        _lib.check(L.sfm_filter_matches_packed(bank.handle, _lib.ptr(pairs_t), P, _lib.ptr(knn), _lib.ptr(rev), C.byref(fprm),
                                               _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(matches), _lib.ptr(corr), st),
                   "sfm_filter_matches_packed")
    total = int(offsets[P].item())
    return counts, offsets, matches[:total], corr[:total]


def match_pairs_hamming(bank: DescriptorBank, pairs, max_distance: int = 26) -> MatchBatch:
    """The reference's literal matcher for every pair: Hamming, crossCheck, sorted by (distance, queryIdx),
    ``distance < max_distance`` (code/feature_matching.py:48-58)."""
    if bank.metric != "hamming":
        raise ValueError("match_pairs_hamming needs a Hamming bank")
    pairs_host = np.asarray(pairs.cpu() if isinstance(pairs, torch.Tensor) else pairs, np.int32).reshape(-1, 2)
    _check_pairs(pairs_host, bank)
    pairs_t = _pairs_tensor(pairs, bank.device)
    P, cap, dev = pairs_t.shape[0], bank.feat_stride, bank.device
    counts = torch.zeros(P, dtype=torch.int32, device=dev)
    matches = torch.empty((P, cap, 3), dtype=torch.int32, device=dev)
    if P == 0:
        return MatchBatch(pairs_t, counts, matches)
    ws = torch.empty(2 * P * cap * 8, dtype=torch.uint8, device=dev)
    _lib.check(
        _lib.lib().sfm_match_hamming(bank.handle, _lib.ptr(pairs_t), P, int(max_distance), _lib.ptr(counts), _lib.ptr(matches),
                                     _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr(dev)),
        "sfm_match_hamming",
    )
    return MatchBatch(pairs_t, counts, matches)


def debug_tc_tile(bank: DescriptorBank, pair, mode: int = 0):
    """Raw tcgen05 accumulators of the first unit / first train tile of one pair (bring-up tests)."""
    pairs_t = _pairs_tensor([pair], bank.device)
    knn = torch.empty((bank.feat_stride, 4), dtype=torch.int32, device=bank.device)
    acc = torch.zeros((256, 128), dtype=torch.int32, device=bank.device)
    _lib.check(
        _lib.lib().sfm_debug_tc_tile(bank.handle, _lib.ptr(pairs_t), int(mode), _lib.ptr(knn), _lib.ptr(acc),
                                     _lib.current_stream_ptr(bank.device)),
        "sfm_debug_tc_tile",
    )
    return acc, knn


def refine_stats(enable: bool = True):
    """(rows brute-forced, candidates recomputed) by the refinement kernel since the counters were switched on."""
    out = (C.c_int64 * 2)()
    _lib.check(_lib.lib().sfm_debug_refine_stats(int(enable), out), "sfm_debug_refine_stats")
    return int(out[0]), int(out[1])


def probe_int8_peak(device: int = 0, n_tiles: int = 4096):
    """(ms, algorithmic int8 op/s) of the epilogue-free MMA loop: the tensor-pipe ceiling of the matcher's tile shape."""
    ms, ops = C.c_float(0), C.c_double(0)
    _lib.check(_lib.lib().sfm_probe_int8_mma(int(device), int(n_tiles), C.byref(ms), C.byref(ops)), "sfm_probe_int8_mma")
    return ms.value, ops.value / (ms.value * 1e-3)
