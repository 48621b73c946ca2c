"""CPU oracle for the descriptor-matching half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product path under
``sfm-project_b200/`` never does (it fails loudly without its CUDA library).

What is restated here, and where it comes from:

* ``hamming_crosscheck`` / ``reference_filter`` follow the reference's literal
  matcher, ``code/feature_matching.py:48-58`` (``cv2.BFMatcher(NORM_HAMMING,
  crossCheck=True).match`` -> ``sorted(key=distance)`` -> prefix ``distance < 26``).
* ``l2_knn2`` / ``ratio_keep`` restate ``cv2.BFMatcher(NORM_L2).knnMatch(k=2)`` plus
  Lowe's ratio idiom, the north-star workload (no call site in the reference;
  SURVEY.md §0 D2/D3/D8).  The arithmetic lives in OpenCV (third party; the
  reference pins no version; this image has opencv-python-headless 4.13.0.92).
  Its observable semantics -- exact integer squared distance, lowest train
  index wins ties, ``distance = float32(sqrt(float32(D)))`` -- are pinned by
  ``tests/test_oracle_pinned.py`` against cv2 itself and against the golden
  vectors in ``tests/golden`` generated from the reference's own function.

Parity status: the reference holds no tests or golden vectors, so
parity is pinned against outputs of the reference function run in the build
container (``tests/golden/make_golden.py``) and against cv2.
"""
from __future__ import annotations

import numpy as np

BIG = np.int64(1) << 40


# --------------------------------------------------------------------------- L2

def sqdist_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Exact int64 squared L2 distances between uint8 rows.

    float32 GEMM is exact here: every partial sum of a.b is an integer
    <= 128*255^2 = 8,323,200 < 2^24.  (Wider rows fall back to float64.)
    """
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    wide = a.shape[1] * 255 * 255 >= (1 << 24)
    ft = np.float64 if wide else np.float32
    dot = (a.astype(ft) @ b.astype(ft).T).astype(np.int64)
    na = (a.astype(np.int64) ** 2).sum(1)
    nb = (b.astype(np.int64) ** 2).sum(1)
    return na[:, None] + nb[None, :] - 2 * dot


def l2_knn2(a: np.ndarray, b: np.ndarray):
    """knnMatch(k=2) restated: (idx1, d1, idx2, d2), int32 each, squared distances.

    Order is (distance, trainIdx) ascending, i.e. the lowest train index wins a
    tie for both neighbours.  Missing neighbours (fewer than 2 train rows) are
    reported as idx=-1, d=-1.
    """
    n1, n2 = a.shape[0], b.shape[0]
    idx1 = np.full(n1, -1, np.int32)
    idx2 = np.full(n1, -1, np.int32)
    d1 = np.full(n1, -1, np.int32)
    d2 = np.full(n1, -1, np.int32)
    if n1 == 0 or n2 == 0:
        return idx1, d1, idx2, d2
    D = sqdist_matrix(a, b)
    rows = np.arange(n1)
    i1 = D.argmin(1)                       # first occurrence == lowest index
    idx1[:] = i1
    d1[:] = D[rows, i1]
    if n2 >= 2:
        D[rows, i1] = BIG
        i2 = D.argmin(1)
        idx2[:] = i2
        d2[:] = D[rows, i2]
    return idx1, d1, idx2, d2


def cv2_distance(d_sq: np.ndarray) -> np.ndarray:
    """DMatch.distance as cv2 reports it for NORM_L2: float32(sqrt(float32(D)))."""
    return np.sqrt(np.asarray(d_sq).astype(np.float32))


def ratio_keep(d1: np.ndarray, d2: np.ndarray, ratio: float = 0.75, mode: str = "cv2_f32") -> np.ndarray:
    """Lowe ratio test on squared distances.

    mode "cv2_f32": the Python idiom ``m.distance < ratio * n.distance`` on cv2's
        float32 distances promoted to Python floats.
    mode "exact_int": ``D1 * den^2 < D2 * num^2`` with ratio = num/den reduced from
        the float (0.75 -> 9*D2 > 16*D1), exact in int64.
    Rows without a second neighbour (d2 < 0) are never kept.
    """
    d1 = np.asarray(d1)
    d2 = np.asarray(d2)
    valid = (d1 >= 0) & (d2 >= 0)
    if mode == "cv2_f32":
        s1 = cv2_distance(np.maximum(d1, 0)).astype(np.float64)
        s2 = cv2_distance(np.maximum(d2, 0)).astype(np.float64)
        keep = s1 < np.float64(ratio) * s2
    elif mode == "exact_int":
        num, den = ratio_as_fraction(ratio)
        keep = d1.astype(np.int64) * (den * den) < d2.astype(np.int64) * (num * num)
    else:
        raise ValueError(f"unknown ratio mode {mode!r}")
    return keep & valid


def ratio_as_fraction(ratio: float):
    from fractions import Fraction

    f = Fraction(ratio).limit_denominator(1024)
    return int(f.numerator), int(f.denominator)


def mutual_mask(idx12: np.ndarray, idx21: np.ndarray) -> np.ndarray:
    """crossCheck: row i survives iff the nearest neighbour of idx12[i] (in the
    reverse direction) is i again."""
    idx12 = np.asarray(idx12)
    ok = idx12 >= 0
    back = np.full(idx12.shape, -2, np.int64)
    back[ok] = np.asarray(idx21)[idx12[ok]]
    return ok & (back == np.arange(idx12.shape[0]))


def match_l2(a, b, *, ratio=0.75, ratio_mode="cv2_f32", mutual=False):
    """Full L2 pair match: returns (q, t, d_sq) int32 arrays in ascending q."""
    idx1, d1, _, d2 = l2_knn2(a, b)
    keep = ratio_keep(d1, d2, ratio, ratio_mode) if ratio is not None else (idx1 >= 0)
    if mutual:
        ridx1, _, _, _ = l2_knn2(b, a)
        keep &= mutual_mask(idx1, ridx1)
    q = np.nonzero(keep)[0].astype(np.int32)
    return q, idx1[q], d1[q]


# ---------------------------------------------------------------------- Hamming

def hamming_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    x = a[:, None, :] ^ b[None, :, :]
    return np.bitwise_count(x).sum(2).astype(np.int32)


def hamming_crosscheck(a: np.ndarray, b: np.ndarray):
    """``BFMatcher(NORM_HAMMING, crossCheck=True).match`` restated
    (code/feature_matching.py:48-50): (q, t, dist) in ascending q."""
    if a is None or b is None or len(a) == 0 or len(b) == 0:
        z = np.zeros(0, np.int32)
        return z, z.copy(), z.copy()
    H = hamming_matrix(a, b)
    nn12 = H.argmin(1)
    nn21 = H.argmin(0)
    rows = np.arange(a.shape[0])
    keep = nn21[nn12] == rows
    q = rows[keep].astype(np.int32)
    t = nn12[keep].astype(np.int32)
    return q, t, H[q, t].astype(np.int32)


def reference_filter(q, t, d, max_distance=26):
    """``sorted(key=distance)`` (stable) then the prefix ``distance < 26``
    (code/feature_matching.py:52-58)."""
    order = np.argsort(d, kind="stable")
    q, t, d = q[order], t[order], d[order]
    n = int(np.searchsorted(d, max_distance, side="left"))
    return q[:n], t[:n], d[:n]


def match_hamming_reference(a, b, max_distance=26):
    """The literal reference pair match on precomputed ORB descriptors."""
    return reference_filter(*hamming_crosscheck(a, b), max_distance=max_distance)
