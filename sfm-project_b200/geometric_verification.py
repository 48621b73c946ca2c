"""Geometric verification: the module the reference left empty.

``code/geometric_verification.py`` is a 0-byte file that ``code/pipeline.py:3`` star-imports, and the
"Geometric Verification" section of ``main()`` is a bare comment (code/pipeline.py:60-65).  This module
defines the API with the conventions of ``cv2.findFundamentalMat(pts1, pts2,
cv2.FM_RANSAC, thr, confidence, maxIters)`` so that cv2 is a one-line oracle: F is float64 [3,3] with
``x2^T F x1 = 0`` and ``F[2,2] == 1``, the mask is uint8 [M,1], and ``(None, zeros)`` is returned when no
model is found.  All computation is the batched CUDA kernel csrc/ransac_f.cu; there is no CPU fallback.
"""
import numpy as np
import torch

import sfm_b200 as _sfm
from sfm_b200.ransac import h_stop_target


def _as_points(pts, name):
    a = np.asarray(pts)
    if a.size == 0:
        return np.zeros((0, 2), np.float32)
    if a.ndim == 3 and a.shape[1] == 1:
        a = a[:, 0, :]
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError(f"{name} must be [M,2] or [M,1,2], got {a.shape}")
    if a.dtype not in (np.float32, np.float64):
        raise ValueError(f"{name} must be float32 or float64")
    return np.ascontiguousarray(a, np.float32)


def verify_pairs(points1, points2, *, thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", score="sym_epipolar",
                 lo=False, seed=0, min_inliers=0):
    """Batched verification.  ``points1[k]``, ``points2[k]`` are the [M_k,2] matched pixel coordinates of pair k.
    Returns a list of ``(F or None, mask uint8 [M_k,1])``."""
    if len(points1) != len(points2):
        raise ValueError("points1 and points2 must have the same number of pairs")
    P = len(points1)
    if P == 0:
        return []
    p1 = [_as_points(p, "pts1") for p in points1]
    p2 = [_as_points(p, "pts2") for p in points2]
    for a, b in zip(p1, p2):
        if len(a) != len(b):
            raise ValueError("pts1 and pts2 must have the same number of points")
    cap = max(16, max(len(a) for a in p1))
    corr = np.zeros((P, cap, 4), np.float32)
    counts = np.zeros(P, np.int32)
    for k, (a, b) in enumerate(zip(p1, p2)):
        corr[k, : len(a), :2], corr[k, : len(a), 2:] = a, b
        counts[k] = len(a)
    vb = _sfm.verify_corr(torch.from_numpy(corr).cuda(), torch.from_numpy(counts), thr=thr, confidence=confidence,
                          max_iters=max_iters, solver=solver, score=score, lo=lo, seed=seed, min_inliers=min_inliers)
    F, ninl, mask = vb.F.cpu().numpy(), vb.n_inliers.cpu().numpy(), vb.mask.cpu().numpy()
    out = []
    for k in range(P):
        m = mask[k, : counts[k]].reshape(-1, 1).copy()
        out.append((F[k].copy() if ninl[k] > 0 else None, m))
    return out


def verify_pair(pts1, pts2, *, thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", score="sym_epipolar", lo=False,
                seed=0):
    """Single-pair form of cv2.findFundamentalMat(FM_RANSAC): returns ``(F float64[3,3] | None, mask uint8[M,1])``."""
    return verify_pairs([pts1], [pts2], thr=thr, confidence=confidence, max_iters=max_iters, solver=solver, score=score,
                        lo=lo, seed=seed)[0]


def verify_matches(kp1, kp2, matches, **kw):
    """Verify a ``list[cv2.DMatch]`` (as returned by ``extract_and_match``) between two keypoint lists.
    Returns ``(F or None, inlier_matches)``."""
    if not matches:
        return None, []
    pts1 = np.array([kp1[m.queryIdx].pt for m in matches], np.float32)
    pts2 = np.array([kp2[m.trainIdx].pt for m in matches], np.float32)
    F, mask = verify_pair(pts1, pts2, **kw)
    if F is None:
        return None, []
    return F, [m for m, keep in zip(matches, mask.ravel()) if keep]


# ------------------------------------------------------------------ homography model
def _corr_batch(points1, points2):
    if len(points1) != len(points2):
        raise ValueError("points1 and points2 must have the same number of pairs")
    p1 = [_as_points(p, "pts1") for p in points1]
    p2 = [_as_points(p, "pts2") for p in points2]
    for a, b in zip(p1, p2):
        if len(a) != len(b):
            raise ValueError("pts1 and pts2 must have the same number of points")
    cap = max(16, max(len(a) for a in p1))
    corr = np.zeros((len(p1), cap, 4), np.float32)
    counts = np.zeros(len(p1), np.int32)
    for k, (a, b) in enumerate(zip(p1, p2)):
        corr[k, : len(a), :2], corr[k, : len(a), 2:] = a, b
        counts[k] = len(a)
    return torch.from_numpy(corr).cuda(), torch.from_numpy(counts), counts


def find_homographies(points1, points2, *, thr=3.0, confidence=0.995, max_iters=2000, lo=False, seed=0, min_inliers=0,
                      stop_targets=None):
    """Batched ``cv2.findHomography(pts1, pts2, cv2.RANSAC, thr)``: a list of ``(H float64[3,3] | None, mask uint8[M_k,1])``
    with ``x2 ~ H x1`` and ``H[2,2] == 1`` (csrc/ransac_h.cu, one CTA per pair).  ``stop_targets[k]`` (optional) is the
    inlier count the caller cares about for pair k: sampling stops once a homography with that support would have been
    found, so a pair that no homography explains costs 32 hypotheses instead of ``max_iters``."""
    if len(points1) == 0:
        return []
    corr, counts_t, counts = _corr_batch(points1, points2)
    vb = _sfm.ransac.verify_h_corr(corr, counts_t, thr=thr, confidence=confidence, max_iters=max_iters, lo=lo, seed=seed,
                                   min_inliers=min_inliers,
                                   stop_target=None if stop_targets is None else np.asarray(stop_targets, np.int32).reshape(len(counts)))
    H, ninl, mask = vb.F.cpu().numpy(), vb.n_inliers.cpu().numpy(), vb.mask.cpu().numpy()
    return [(H[k].copy() if ninl[k] > 0 else None, mask[k, : counts[k]].reshape(-1, 1).copy()) for k in range(len(counts))]


def find_homography(pts1, pts2, *, stop_target=None, **kw):
    """Single-pair form: ``(H | None, mask uint8[M,1])``."""
    return find_homographies([pts1], [pts2], stop_targets=None if stop_target is None else [stop_target], **kw)[0]


# ------------------------------------------------------------------ two-view initialisation
def recover_poses(Fs, points1, points2, K1, K2=None, *, masks=None, distance_thresh=50.0):
    """Batched ``cv2.recoverPose`` starting from F: ``E = K2^T F K1``, the (R, t) among the four decompositions with
    the most correspondences in front of both cameras, and their DLT triangulation (csrc/pose.cu).  ``K1`` / ``K2``
    are 3x3 pinhole matrices (shared) or ``[P,3,3]`` stacks (``K2=None``: same camera).  Returns a list of
    ``(n_good, R float64[3,3], t float64[3], mask uint8[M_k,1], points3d float32[M_k,3])``; points are expressed in
    the first camera's frame and are zero where the mask is 0; ``n_good == 0`` means no pose."""
    P = len(points1)
    if P == 0:
        return []
    corr, counts_t, counts = _corr_batch(points1, points2)
    F = np.zeros((P, 3, 3), np.float64)
    for k, f in enumerate(Fs):
        if f is not None:
            f = np.asarray(f, np.float64)
            if f.shape != (3, 3):
                raise ValueError(f"F must be [3,3], got {f.shape}")
            F[k] = f
    m = None
    if masks is not None:
        m = torch.zeros((P, corr.shape[1]), dtype=torch.uint8)
        for k, mk in enumerate(masks):
            mk = np.asarray(mk).reshape(-1)
            if len(mk) != counts[k]:
                raise ValueError("mask length must equal the number of points of its pair")
            m[k, : counts[k]] = torch.from_numpy((mk != 0).astype(np.uint8))
    cam = _sfm.ransac.camera_rows(K1, K2, P)
    pb = _sfm.ransac.recover_pose_corr(corr, counts_t, torch.from_numpy(F), cam, mask=m, distance_thresh=distance_thresh)
    R, t, ng = pb.R.cpu().numpy(), pb.t.cpu().numpy(), pb.n_good.cpu().numpy()
    pm, X = pb.mask.cpu().numpy(), pb.points.cpu().numpy()
    return [(int(ng[k]), R[k].copy(), t[k].copy(), pm[k, : counts[k]].reshape(-1, 1).copy(), X[k, : counts[k]].copy()) for k in range(P)]


def recover_pose(F, pts1, pts2, K1, K2=None, *, mask=None, distance_thresh=50.0):
    """Single-pair form: ``(n_good, R, t, mask, points3d)``."""
    return recover_poses([F], [pts1], [pts2], K1, K2, masks=None if mask is None else [mask], distance_thresh=distance_thresh)[0]


# ------------------------------------------------------------------ scene-graph classification
DEGENERATE, PLANAR_OR_PANORAMIC, UNCALIBRATED, CALIBRATED = "degenerate", "planar_or_panoramic", "uncalibrated", "calibrated"


def classify_pairs(n_inliers_f, n_inliers_h, *, min_inliers=15, max_h_inlier_ratio=0.8, calibrated=False):
    """Scene-graph edge type per pair from the two inlier counts (the two-view geometry test of Schoenberger & Frahm
    2016 §4.1, ``papers/schoenberger2016sfm.pdf`` in the reference): a pair is kept iff F explains at least
    ``min_inliers`` matches; it is planar / panoramic (not usable to seed triangulation) when H explains more than
    ``max_h_inlier_ratio`` of what F explains."""
    nf = np.asarray(n_inliers_f, np.int64)
    nh = np.asarray(n_inliers_h, np.int64)
    out = np.full(nf.shape, CALIBRATED if calibrated else UNCALIBRATED, dtype=object)
    out[nh > max_h_inlier_ratio * nf] = PLANAR_OR_PANORAMIC
    out[nf < min_inliers] = DEGENERATE
    return out


def two_view_geometry(pts1, pts2, K1=None, K2=None, *, thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", lo=True, seed=0,
                      min_inliers=15, max_h_inlier_ratio=0.8, distance_thresh=50.0):
    """Everything the scene graph stores for one pair: F and its inliers, H and its inliers, the edge type, and with
    intrinsics the relative pose and the triangulated inliers.  All three estimators run on the GPU."""
    F, fmask = verify_pair(pts1, pts2, thr=thr, confidence=confidence, max_iters=max_iters, solver=solver, lo=lo, seed=seed)
    H, hmask = find_homography(pts1, pts2, thr=thr, confidence=confidence, max_iters=max_iters, lo=lo, seed=seed,
                               stop_target=h_stop_target(int(fmask.sum()), max_h_inlier_ratio))
    out = {"F": F, "inlier_mask": fmask, "n_inliers": int(fmask.sum()), "H": H, "inlier_mask_h": hmask, "n_inliers_h": int(hmask.sum())}
    out["config"] = classify_pairs([out["n_inliers"]], [out["n_inliers_h"]], min_inliers=min_inliers,
                                   max_h_inlier_ratio=max_h_inlier_ratio, calibrated=K1 is not None)[0]
    if K1 is not None and F is not None:
        n, R, t, pm, X = recover_pose(F, pts1, pts2, K1, K2, mask=fmask, distance_thresh=distance_thresh)
        out.update(n_pose=n, R=R if n else None, t=t if n else None, in_front=pm, points3d=X)
    return out
