"""Execution plan of the hot path for batches of image pairs: preallocated device and pinned-host buffers, a
fixed launch sequence on the caller's stream, and result copies on a side stream that overlap the RANSAC kernel.

    sweep (tcgen05) -> refine -> [reverse sweep -> refine]      sfm_match_knn2
    count -> scan -> write (packed matches + correspondences)    sfm_filter_matches_packed
    RANSAC-F on the packed correspondences                       sfm_ransac_f_packed

This is the batched body of the reference's pair loop (code/pipeline.py:38-47) plus the verification stage it left
empty (code/pipeline.py:60-65).  Nothing here computes: every array is produced by lib/libsfm_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .bank import DescriptorBank
from .matcher import filter_params
from .ransac import ransac_params


class HotPathPlan:
    """Buffers + launch sequence for up to ``max_pairs`` pairs per batch on one bank.

    Packed result layout (device and host): pair p of the batch owns rows [offsets[p], offsets[p+1]) of
    ``matches`` int32 [total,3] = (queryIdx, trainIdx, squared L2), ``corr`` float32 [total,4] and ``mask`` uint8 [total].
    """

    def __init__(self, bank: DescriptorBank, max_pairs: int, *, ratio=0.75, ratio_mode="cv2_f32", mutual=False, impl="auto",
                 thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", score="sym_epipolar", lo=False, seed=0,
                 min_inliers=0, prefilter=True):
        if bank.metric != "l2":
            raise ValueError("the verification path needs an L2 bank")
        self.bank, self.B, self.cap, self.dev = bank, int(max_pairs), bank.feat_stride, bank.device
        if self.B < 1:
            raise ValueError("max_pairs must be positive")
        if self.B * self.cap >= 2 ** 31:
            raise ValueError("max_pairs * feat_stride must stay below 2^31 (int32 offsets); use smaller batches")
        self.mutual = bool(mutual)
        self.fprm = filter_params(ratio, ratio_mode, mutual)
        self.rprm = ransac_params(thr=thr, confidence=confidence, max_iters=max_iters, solver=solver, score=score, lo=lo,
                                  seed=seed, min_inliers=min_inliers)
        self.mprm = _lib.MatchParams()
        self.mprm.impl = _lib.MATCH_IMPLS[impl]
        self.prefilter = bool(prefilter) and not self.mutual
        B, cap, dev = self.B, self.cap, self.dev
        i32 = dict(dtype=torch.int32, device=dev)
        self.knn = torch.empty((B, cap, 4), **i32)
        self.knn_rev = torch.empty((B, cap, 4), **i32) if self.mutual else None
        self.counts = torch.zeros(B, **i32)
        self.offsets = torch.zeros(B + 1, **i32)
        self.matches = torch.empty((B * cap, 3), **i32)
        self.corr = torch.empty((B * cap, 4), dtype=torch.float32, device=dev)
        self.mask = torch.empty(B * cap, dtype=torch.uint8, device=dev)
        self.F = torch.zeros((B, 3, 3), dtype=torch.float64, device=dev)
        self.ninl = torch.zeros(B, **i32)
        self.iters = torch.zeros(B, **i32)
        # pinned host side
        self.offsets_h = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
        self.F_h = torch.zeros((B, 3, 3), dtype=torch.float64).pin_memory()
        self.ninl_h = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.iters_h = torch.zeros(B, dtype=torch.int32).pin_memory()
        self._rows_h = 0
        self.matches_h = self.mask_h = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_filter = torch.cuda.Event()
        self.ev_done = torch.cuda.Event()
        self.ev_offsets = torch.cuda.Event()
        self.ev_copied = torch.cuda.Event()
        self._copy_pending = False
        self.P = self._fetch_P = self._fetch_total = 0

    # ------------------------------------------------------------------ launches (no host synchronisation)
    def launch(self, pairs_d: torch.Tensor, pair_id_d: torch.Tensor, pairs_rev_d: torch.Tensor | None = None) -> int:
        """Enqueue match -> filter -> verify on the current stream for ``pairs_d`` int32 [P,2] (device, image ids already
        validated by the caller), ``pair_id_d`` int32 [P] (RANSAC stream ids) and, for mutual matching, the swapped pair
        list.  Returns P."""
        P = int(pairs_d.shape[0])
        if P > self.B:
            raise ValueError(f"{P} pairs exceed the plan's batch size {self.B}")
        self.P = P
        self._last_pair_id = pair_id_d
        if P == 0:
            return 0
        L, st, bank = _lib.lib(), _lib.current_stream_ptr(self.dev), self.bank
        cur = torch.cuda.current_stream(self.dev)
        if self.prefilter:
            # the sweep marks rows that cannot pass the ratio test; the refinement skips them (they read as "no match")
            self.mprm.prefilter_mode = self.fprm.ratio_mode
            self.mprm.prefilter_ratio = self.fprm.ratio
            self.mprm.prefilter_num, self.mprm.prefilter_den = int(self.fprm.ratio_num), int(self.fprm.ratio_den)
        _lib.check(L.sfm_match_knn2(bank.handle, _lib.ptr(pairs_d), P, C.byref(self.mprm), _lib.ptr(self.knn), None, 0, st),
                   "sfm_match_knn2")
        if self.mutual:
            if pairs_rev_d is None:
                pairs_rev_d = pairs_d.flip(1).contiguous()
            plain = _lib.MatchParams()
            plain.impl = self.mprm.impl
            _lib.check(L.sfm_match_knn2(bank.handle, _lib.ptr(pairs_rev_d), P, C.byref(plain), _lib.ptr(self.knn_rev), None, 0, st),
                       "sfm_match_knn2 (reverse)")
        if self._copy_pending:                       # the previous batch's result copies still read the packed buffers
            cur.wait_event(self.ev_copied)
            self._copy_pending = False
        _lib.check(L.sfm_filter_matches_packed(bank.handle, _lib.ptr(pairs_d), P, _lib.ptr(self.knn), _lib.ptr(self.knn_rev),
                                               C.byref(self.fprm), _lib.ptr(self.counts), _lib.ptr(self.offsets),
                                               _lib.ptr(self.matches), _lib.ptr(self.corr), st), "sfm_filter_matches_packed")
        self.ev_filter.record(cur)
        _lib.check(L.sfm_ransac_f_packed(_lib.ptr(self.corr), _lib.ptr(self.offsets), P, self.cap, _lib.ptr(pair_id_d), None,
                                         C.byref(self.rprm), _lib.ptr(self.F), _lib.ptr(self.ninl), _lib.ptr(self.mask),
                                         _lib.ptr(self.iters), st), "sfm_ransac_f_packed")
        self.ev_done.record(cur)
        return P

    def rerun_ransac(self) -> None:
        """Verification stage alone on the packed correspondences of the last batch (bench.py times it in isolation)."""
        if self.P == 0:
            return
        _lib.check(_lib.lib().sfm_ransac_f_packed(_lib.ptr(self.corr), _lib.ptr(self.offsets), self.P, self.cap,
                                                  _lib.ptr(self._last_pair_id), None, C.byref(self.rprm), _lib.ptr(self.F),
                                                  _lib.ptr(self.ninl), _lib.ptr(self.mask), _lib.ptr(self.iters),
                                                  _lib.current_stream_ptr(self.dev)), "sfm_ransac_f_packed")

    # ------------------------------------------------------------------ results
    def _ensure_rows(self, rows: int) -> None:
        if rows > self._rows_h:
            rows = max(rows, 2 * self._rows_h, 1 << 16)
            self.matches_h = torch.empty((rows, 3), dtype=torch.int32).pin_memory()
            self.mask_h = torch.empty(rows, dtype=torch.uint8).pin_memory()
            self._rows_h = rows

    def fetch_begin(self) -> None:
        """Enqueue the device -> pinned-host copies of the last batch on the side stream.  Blocks the host only until
        the filter has finished (to learn the packed size); the match rows then travel while RANSAC is running."""
        P = self._fetch_P = self.P
        if P == 0:
            return
        cs = self.copy_stream
        with torch.cuda.stream(cs):
            cs.wait_event(self.ev_filter)
            self.offsets_h[: P + 1].copy_(self.offsets[: P + 1], non_blocking=True)
            self.ev_offsets.record(cs)
        self.ev_offsets.synchronize()                 # filter finished; RANSAC keeps the GPU busy meanwhile
        total = self._fetch_total = int(self.offsets_h[P])
        self._ensure_rows(total)
        with torch.cuda.stream(cs):
            if total:
                self.matches_h[:total].copy_(self.matches[:total], non_blocking=True)
            cs.wait_event(self.ev_done)
            if total:
                self.mask_h[:total].copy_(self.mask[:total], non_blocking=True)
            self.F_h[:P].copy_(self.F[:P], non_blocking=True)
            self.ninl_h[:P].copy_(self.ninl[:P], non_blocking=True)
            self.iters_h[:P].copy_(self.iters[:P], non_blocking=True)
            self.ev_copied.record(cs)
        self._copy_pending = True

    def fetch_end(self) -> dict:
        """Wait for the copies of ``fetch_begin`` and return numpy VIEWS of the pinned buffers (valid until the next
        ``fetch_begin`` on this plan)."""
        P = self._fetch_P
        if P == 0:
            return {"n_matches": np.zeros(0, np.int32), "offsets": np.zeros(1, np.int64), "matches": np.zeros((0, 3), np.int32),
                    "inlier": np.zeros(0, np.uint8), "F": np.zeros((0, 3, 3)), "n_inliers": np.zeros(0, np.int32),
                    "iters": np.zeros(0, np.int32)}
        total = self._fetch_total
        self.ev_copied.synchronize()
        off = self.offsets_h[: P + 1].numpy()
        return {"n_matches": np.diff(off).astype(np.int32), "offsets": off.astype(np.int64),
                "matches": self.matches_h[:total].numpy(), "inlier": self.mask_h[:total].numpy(),
                "F": self.F_h[:P].numpy(), "n_inliers": self.ninl_h[:P].numpy(), "iters": self.iters_h[:P].numpy()}

    def fetch(self) -> dict:
        self.fetch_begin()
        return self.fetch_end()

    def host_bytes(self) -> int:
        """Bytes the last fetch moved device -> host."""
        P = self._fetch_P
        return (4 * (P + 1) + self._fetch_total * 13 + P * (72 + 4 + 4)) if P else 0
