"""The reference's real arithmetic: OpenCV calls.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's hot path is a handful of cv2 calls (code/feature_matching.py:42-58);
the north-star workload swaps in BFMatcher(NORM_L2).knnMatch + findFundamentalMat
(BASELINE.json configs[0]).  cv2 (opencv-python-headless 4.13.0) is present in this
image and on the GPU box, so these wrappers are used (a) to pin the numpy / C
restatements in oracle/, (b) as the ``--impl reference`` arm and ``cpu_baseline``
of bench.py.  Nothing under sfm-project_b200/ imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

import cv2

REFERENCE_CODE_DIR = "/root/reference/code"


def import_reference_feature_matching():
    """Import the reference's own ``feature_matching`` module (build container only;
    /root/reference does not exist on the GPU box).  ``matplotlib`` is not installed,
    so a stub satisfies ``import matplotlib.pyplot as plt`` (code/feature_matching.py:5).
    The module is loaded under a private name so it never shadows the drop-in."""
    path = os.path.join(REFERENCE_CODE_DIR, "feature_matching.py")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    spec = importlib.util.spec_from_file_location("_reference_feature_matching", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dmatches_to_arrays(matches):
    q = np.array([m.queryIdx for m in matches], np.int32)
    t = np.array([m.trainIdx for m in matches], np.int32)
    d = np.array([m.distance for m in matches], np.float32)
    return q, t, d


def orb_extract(gray):
    """The extraction step the reference runs per pair (code/feature_matching.py:42-45)."""
    orb = cv2.ORB_create()
    return orb.detectAndCompute(gray, None)


def hamming_crosscheck(des1, des2):
    """code/feature_matching.py:48-50 verbatim on descriptors."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    return dmatches_to_arrays(bf.match(des1, des2))


def l2_knn2(des1, des2):
    """BFMatcher(NORM_L2).knnMatch(k=2) on float32 copies (3.6x faster than uint8 in cv2,
    identical results: SURVEY.md A.1).  Returns idx1, dist1(f32), idx2, dist2(f32)."""
    a = np.ascontiguousarray(des1, np.float32)
    b = np.ascontiguousarray(des2, np.float32)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(a, b, k=2)
    n = len(knn)
    idx1 = np.full(n, -1, np.int32)
    idx2 = np.full(n, -1, np.int32)
    d1 = np.full(n, -1, np.float32)
    d2 = np.full(n, -1, np.float32)
    for r, ms in enumerate(knn):
        if len(ms) > 0:
            idx1[r], d1[r] = ms[0].trainIdx, ms[0].distance
        if len(ms) > 1:
            idx2[r], d2[r] = ms[1].trainIdx, ms[1].distance
    return idx1, d1, idx2, d2


def l2_ratio_match(des1, des2, ratio=0.75):
    """knnMatch + the ubiquitous ``m.distance < ratio * n.distance`` loop.  Returns (q, t, dist)."""
    a = np.ascontiguousarray(des1, np.float32)
    b = np.ascontiguousarray(des2, np.float32)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(a, b, k=2)
    good = []
    for ms in knn:
        if len(ms) == 2 and ms[0].distance < ratio * ms[1].distance:
            good.append(ms[0])
    return dmatches_to_arrays(good)


def l2_crosscheck(des1, des2):
    a = np.ascontiguousarray(des1, np.float32)
    b = np.ascontiguousarray(des2, np.float32)
    return dmatches_to_arrays(cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(a, b))


def find_fundamental(pts1, pts2, thr=3.0, confidence=0.99, max_iters=2000):
    """cv2.findFundamentalMat(FM_RANSAC): 7-point samples, sym-epipolar-max metric, no refit
    (SURVEY.md D6).  Returns (F or None, mask uint8[M])."""
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    if len(p1) < 8:
        return None, np.zeros(len(p1), np.uint8)
    F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, thr, confidence, max_iters)
    if F is None or F.shape != (3, 3):
        return None, np.zeros(len(p1), np.uint8)
    return F, mask.ravel().astype(np.uint8)


def verified_pair(des1, xy1, des2, xy2, *, ratio=0.75, thr=3.0, confidence=0.99, max_iters=2000):
    """One 'verified pair' of BASELINE.json configs[0]: knnMatch + ratio + RANSAC-F."""
    q, t, _ = l2_ratio_match(des1, des2, ratio)
    F, mask = find_fundamental(xy1[q], xy2[t], thr, confidence, max_iters)
    return q, t, F, mask
