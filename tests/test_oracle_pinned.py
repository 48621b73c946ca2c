"""Pin the CPU oracles: against the committed golden fixtures (outputs of the reference's own
function and of cv2, tests/golden/make_golden.py) and against cv2 run live.  CPU only."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import match_oracle as mo
from oracle import ransac_oracle as ro
from sfm_b200 import synth

cv2 = pytest.importorskip("cv2")
from oracle import cv2_ref  # noqa: E402


# ------------------------------------------------------------------ golden: reference function
def test_hamming_oracle_matches_reference_function_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    n = g["images"].shape[0]
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            q, t, d = mo.match_hamming_reference(g[f"des{i}"], g[f"des{j}"], 26)
            assert q.tolist() == g[f"q_{i}_{j}"].tolist()
            assert t.tolist() == g[f"t_{i}_{j}"].tolist()
            assert d.astype(np.float32).tolist() == g[f"d_{i}_{j}"].tolist()
            assert (d < 26).all() and len(q) > 50


def test_golden_orb_descriptors_reproduce(golden_dir):
    """cv2 ORB on the stored images gives the stored descriptors (extraction is deterministic),
    so the GPU drop-in test can start from images."""
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    kp, des = cv2_ref.orb_extract(g["images"][0])
    assert np.array_equal(des, g["des0"])


# ------------------------------------------------------------------ golden: cv2 L2 knn
def test_l2_knn_oracle_matches_cv2_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_l2_knn.npz"))
    idx1, d1, idx2, d2 = mo.l2_knn2(g["A"], g["B"])
    assert idx1.tolist() == g["idx1"].tolist()
    assert idx2.tolist() == g["idx2"].tolist()
    assert mo.cv2_distance(d1).tolist() == g["d1"].tolist()      # bitwise float32
    assert mo.cv2_distance(d2).tolist() == g["d2"].tolist()
    # triplicated row: lowest indices first
    assert (idx1[200], idx2[200]) == (5, 300) and d1[200] == 0 and d2[200] == 0
    keep = mo.ratio_keep(d1, d2, 0.75, "cv2_f32")
    assert np.nonzero(keep)[0].tolist() == g["ratio_q"].tolist()
    assert idx1[keep].tolist() == g["ratio_t"].tolist()
    # documented divergence of the exact-integer ratio test at 16*D1 == 9*D2
    assert (d1[202], d2[202]) == (18, 32) and (d1[203], d2[203]) == (27, 48)
    ki = mo.ratio_keep(d1, d2, 0.75, "exact_int")
    assert keep[202] and not ki[202]
    assert keep[203] and not ki[203]
    other = np.ones(len(keep), bool)
    other[[202, 203]] = False
    assert np.array_equal(keep[other], ki[other])


def test_l2_mutual_oracle_matches_cv2_crosscheck_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_l2_knn.npz"))
    q, t, d = mo.match_l2(g["A"], g["B"], ratio=None, mutual=True)
    assert q.tolist() == g["cross_q"].tolist()
    assert t.tolist() == g["cross_t"].tolist()
    assert mo.cv2_distance(d).tolist() == g["cross_d"].tolist()


# ------------------------------------------------------------------ live cv2
@pytest.mark.parametrize("seed,n1,n2", [(0, 257, 300), (1, 64, 1), (2, 1, 50), (3, 500, 333)])
def test_l2_knn_oracle_vs_cv2_live(seed, n1, n2):
    rng = np.random.default_rng(seed)
    B = synth.sift_like(rng, n2)
    A = synth.sift_like(rng, n1)
    k = min(n1, n2) // 2
    if k:
        A[:k] = synth.observe(rng, B[rng.permutation(n2)[:k]])
    c1, cd1, c2, cd2 = cv2_ref.l2_knn2(A, B)
    idx1, d1, idx2, d2 = mo.l2_knn2(A, B)
    assert idx1.tolist() == c1.tolist() and idx2.tolist() == c2.tolist()
    assert mo.cv2_distance(d1).tolist() == cd1.tolist()
    if n2 >= 2:
        assert mo.cv2_distance(d2).tolist() == cd2.tolist()
    else:
        assert (idx2 == -1).all()


def test_hamming_oracle_vs_cv2_live_with_ties():
    rng = np.random.default_rng(5)
    for trial in range(6):
        n1, n2 = rng.integers(50, 400, 2)
        b = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        a = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        k = min(n1, n2) // 2
        a[:k] = b[rng.permutation(n2)[:k]] ^ (rng.random((k, 32)) < 0.02).astype(np.uint8)
        a[k // 2] = a[0]                      # duplicate query rows
        b[n2 - 1] = b[0]                      # duplicate train rows
        cq, ct, cd = cv2_ref.hamming_crosscheck(a, b)
        q, t, d = mo.hamming_crosscheck(a, b)
        assert q.tolist() == cq.tolist() and t.tolist() == ct.tolist()
        assert d.astype(np.float32).tolist() == cd.tolist()


def test_hamming_empty_sides():
    a = np.zeros((0, 32), np.uint8)
    b = np.ones((4, 32), np.uint8)
    assert len(mo.match_hamming_reference(a, b)[0]) == 0
    assert len(mo.match_hamming_reference(b, a)[0]) == 0
    assert len(mo.match_hamming_reference(None, b)[0]) == 0


# ------------------------------------------------------------------ RANSAC oracle
def test_fm_metric_reproduces_cv2_mask_golden(golden_dir):
    """cv2's returned mask == max(d1^2,d2^2) <= thr^2 evaluated on cv2's returned F, both in the
    float64 restatement and in the oracle's float32 scorer."""
    g = np.load(os.path.join(golden_dir, "cv2_fm_ransac.npz"))
    e = ro.sym_epipolar_err(g["F"], g["pts1"], g["pts2"])
    assert ((e <= 9.0).astype(np.uint8) == g["mask"]).all()
    n, m = ro.count_inliers(g["F"] / np.linalg.norm(g["F"]), g["pts1"], g["pts2"], 3.0, 0)
    far = np.abs(e - 9.0) > 1e-3 * 9.0            # float32 scoring may flip only borderline points
    assert (m[far] == g["mask"][far]).all()
    assert abs(n - int(g["mask"].sum())) <= int((~far).sum())


def test_minimal_solvers_exact_data():
    p1, p2, _, Ft = synth.two_view_correspondences(64, outlier_frac=0.0, seed=5, pixel_sigma=0.0)
    for m in (7, 8):
        Fs = ro.solve_minimal(p1, p2, np.arange(m) * 3)
        assert 1 <= len(Fs) <= 3
        best = min(ro.sym_epipolar_err(F, p1, p2).max() for F in Fs)
        assert best < 1e-4                             # px^2 (float32 input rounding only)
        for F in Fs:
            assert abs(np.linalg.det(F)) < 1e-12
            assert ro.sym_epipolar_err(F, p1[np.arange(m) * 3], p2[np.arange(m) * 3]).max() < 1e-6


def test_minimal_solver_agrees_with_numpy_svd():
    p1, p2, _, _ = synth.two_view_correspondences(40, outlier_frac=0.0, seed=9, pixel_sigma=0.5)
    idx = np.array([1, 5, 9, 13, 17, 21, 25, 29])
    F = ro.solve_minimal(p1, p2, idx)[0]
    # independent float64 8-point on the same sample (no normalisation needed at this size)
    x1, x2 = p1[idx].astype(np.float64), p2[idx].astype(np.float64)
    c1, c2 = x1.mean(0), x2.mean(0)
    s1 = np.sqrt(2) / np.linalg.norm(x1 - c1, axis=1).mean()
    s2 = np.sqrt(2) / np.linalg.norm(x2 - c2, axis=1).mean()
    u1, u2 = (x1 - c1) * s1, (x2 - c2) * s2
    A = np.stack([u2[:, 0] * u1[:, 0], u2[:, 0] * u1[:, 1], u2[:, 0], u2[:, 1] * u1[:, 0], u2[:, 1] * u1[:, 1], u2[:, 1],
                  u1[:, 0], u1[:, 1], np.ones(8)], 1)
    f = np.linalg.svd(A)[2][-1].reshape(3, 3)
    U, S, Vt = np.linalg.svd(f)
    f = U @ np.diag([S[0], S[1], 0]) @ Vt
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1]])
    Fn = T2.T @ f @ T1
    Fn /= np.linalg.norm(Fn)
    if np.sign(Fn.ravel()[np.abs(Fn).argmax()]) != np.sign(F.ravel()[np.abs(Fn).argmax()]):
        Fn = -Fn
    assert np.abs(Fn - F).max() < 1e-9


@pytest.mark.parametrize("outl,n,solver", [(0.3, 1000, 7), (0.5, 2000, 7), (0.5, 2000, 8)])
def test_ransac_oracle_iou_vs_ground_truth_at_least_cv2(outl, n, solver):
    """SURVEY D7 gate: IoU against synthetic ground truth >= cv2's IoU against the same truth."""
    ious, cious = [], []
    for seed in range(3):
        p1, p2, gt, _ = synth.two_view_correspondences(n, outlier_frac=outl, seed=100 + seed)
        F, m, ninl, iters = ro.ransac_f(p1, p2, solver=solver, thr=3.0, max_iters=2000, confidence=0.99, seed=seed, lo=True)
        Fc, mc = cv2_ref.find_fundamental(p1, p2, 3.0, 0.99, 2000)
        assert F is not None and ninl == int(m.sum()) and iters % 32 == 0 or iters == 2000
        assert abs(F[2, 2] - 1.0) < 1e-12
        # returned F: Sampson residual of ground-truth inliers is small
        assert np.median(ro.sampson_err(F, p1[gt], p2[gt])) < 1.0
        ious.append(ro.iou(m, gt))
        cious.append(ro.iou(mc, gt))
    assert np.mean(ious) >= np.mean(cious) - 0.005
    assert np.mean(ious) > 0.97


def test_ransac_oracle_degenerate_inputs():
    p1, p2, _, _ = synth.two_view_correspondences(6, outlier_frac=0.0, seed=1)
    F, m, n, it = ro.ransac_f(p1, p2, solver=7)
    assert F is None and n == 0 and m.sum() == 0 and it == 0
    F, m, n, it = ro.ransac_f(p1[:0], p2[:0], solver=8)
    assert F is None and n == 0
    same = np.tile(np.array([[10.0, 20.0]], np.float32), (50, 1))
    F, m, n, it = ro.ransac_f(same, same, solver=7, max_iters=256)
    assert F is None and n == 0           # coincident points: no model, no crash


def test_ransac_oracle_explicit_samples_and_determinism():
    p1, p2, gt, _ = synth.two_view_correspondences(300, outlier_frac=0.2, seed=4)
    rng = np.random.default_rng(0)
    samples = rng.integers(0, 300, (256, 8)).astype(np.uint32)
    r1 = ro.ransac_f(p1, p2, solver=8, max_iters=256, confidence=1.0, samples=samples)
    r2 = ro.ransac_f(p1, p2, solver=8, max_iters=256, confidence=1.0, samples=samples)
    assert r1[2] == r2[2] and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[0], r2[0])
    assert r1[3] == 256
    s = ro.draw_sample(7, 3, 11, 8, 300)
    assert len(set(s.tolist())) == 8 and s.min() >= 0 and s.max() < 300
    assert ro.draw_sample(7, 3, 11, 8, 300).tolist() == s.tolist()


# ------------------------------------------------------------------ homography oracle
def test_homography_metric_reproduces_cv2_mask_golden(golden_dir):
    """cv2.findHomography's returned mask == forward transfer error <= thr^2 on cv2's returned H, in the float64
    restatement and (away from the boundary) in the oracle's float32 division-free scorer."""
    g = np.load(os.path.join(golden_dir, "cv2_two_view.npz"))
    e = ro.transfer_err(g["H"], g["h_pts1"], g["h_pts2"])
    assert ((e <= 9.0).astype(np.uint8) == g["h_mask"]).all()
    corr = np.concatenate([g["h_pts1"], g["h_pts2"]], 1).astype(np.float32)
    m = np.zeros(len(corr), np.uint8)
    Hn = np.ascontiguousarray(g["H"] / np.linalg.norm(g["H"]))
    import ctypes as C
    n = ro.lib().sfm_oracle_count_inliers_h(Hn.ctypes.data_as(C.c_void_p), corr.ctypes.data_as(C.c_void_p), C.c_int(len(corr)),
                                            C.c_float(3.0), m.ctypes.data_as(C.c_void_p))
    far = np.abs(e - 9.0) > 1e-3 * 9.0
    assert (m[far] == g["h_mask"][far]).all() and abs(n - int(g["h_mask"].sum())) <= int((~far).sum())


def test_homography_minimal_solver_exact_and_vs_numpy():
    p1, p2, _, Ht = synth.planar_correspondences(64, outlier_frac=0.0, seed=5, pixel_sigma=0.0)
    idx = np.array([3, 17, 40, 58])
    H = ro.solve_h4(p1, p2, idx)
    assert H is not None and abs(np.linalg.norm(H) - 1.0) < 1e-12
    assert ro.transfer_err(H, p1, p2).max() < 1e-4                 # px^2: float32 input rounding only
    assert np.abs(H / H[2, 2] - Ht).max() < 1e-3 * np.abs(Ht).max()
    # independent float64 DLT by SVD on the same (noisy) sample
    p1, p2, _, _ = synth.planar_correspondences(64, outlier_frac=0.0, seed=6, pixel_sigma=0.5)
    H = ro.solve_h4(p1, p2, idx)
    x1, x2 = p1[idx].astype(np.float64), p2[idx].astype(np.float64)
    A = []
    for (a, b), (c, d) in zip(x1, x2):
        A.append([a, b, 1, 0, 0, 0, -c * a, -c * b, -c])
        A.append([0, 0, 0, a, b, 1, -d * a, -d * b, -d])
    Hs = np.linalg.svd(np.asarray(A))[2][-1].reshape(3, 3)
    Hs = Hs / np.linalg.norm(Hs) * np.sign(Hs[2, 2]) * np.sign(H[2, 2])
    assert np.abs(Hs - H).max() < 1e-9
    # three collinear sample points: no model
    q1 = p1.copy()
    q1[idx[:3]] = np.array([[10, 10], [20, 20], [30, 30]], np.float32)
    assert ro.solve_h4(q1, p2, idx) is None


@pytest.mark.parametrize("outl,n", [(0.3, 1000), (0.5, 2000)])
def test_homography_oracle_iou_vs_ground_truth_at_least_cv2(outl, n):
    ious, cious = [], []
    for seed in range(3):
        p1, p2, gt, Ht = synth.planar_correspondences(n, outlier_frac=outl, seed=200 + seed)
        H, m, ninl, iters = ro.ransac_h(p1, p2, thr=3.0, max_iters=2000, confidence=0.995, seed=seed, lo=True)
        Hc, mc = cv2.findHomography(p1, p2, cv2.RANSAC, 3.0, maxIters=2000, confidence=0.995)
        assert H is not None and ninl == int(m.sum()) and H[2, 2] == 1.0
        assert np.median(ro.transfer_err(H, p1[gt], p2[gt])) < 1.5
        ious.append(ro.iou(m, gt))
        cious.append(ro.iou(mc.ravel(), gt))
    assert np.mean(ious) >= np.mean(cious) - 0.005 and np.mean(ious) > 0.97


def test_homography_oracle_degenerate_and_determinism():
    p1, p2, _, _ = synth.planar_correspondences(3, outlier_frac=0.0, seed=1)
    H, m, n, it = ro.ransac_h(p1, p2)
    assert H is None and n == 0 and it == 0
    same = np.tile(np.array([[10.0, 20.0]], np.float32), (50, 1))
    H, m, n, it = ro.ransac_h(same, same, max_iters=256)
    assert H is None and n == 0 and m.sum() == 0
    p1, p2, _, _ = synth.planar_correspondences(300, outlier_frac=0.2, seed=4)
    samples = np.random.default_rng(0).integers(0, 300, (256, 8)).astype(np.uint32)
    r1 = ro.ransac_h(p1, p2, max_iters=256, confidence=1.0, samples=samples)
    r2 = ro.ransac_h(p1, p2, max_iters=256, confidence=1.0, samples=samples)
    assert r1[3] == 256 and r1[2] == r2[2] and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[0], r2[0])
    # a general (non-planar) scene is NOT explained by a homography: far fewer inliers than F finds
    q1, q2, gt, _ = synth.two_view_correspondences(1000, outlier_frac=0.2, seed=3)
    assert ro.ransac_h(q1, q2, seed=1)[2] < 0.5 * ro.ransac_f(q1, q2, solver=7, seed=1)[2]


# ------------------------------------------------------------------ two-view pose oracle
def _cam8(K):
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]] * 2, np.float64)


def test_pose_oracle_matches_cv2_recover_pose_golden(golden_dir):
    """Tolerances: R, t within 1e-9 of cv2.recoverPose (observed 1e-15), cheirality mask identical, triangulated points
    within 1e-5 relative of cv2.triangulatePoints (float32 output)."""
    g = np.load(os.path.join(golden_dir, "cv2_two_view.npz"))
    n, R, t, E, pm, X = ro.two_view_pose(g["pts1"], g["pts2"], g["F"], _cam8(g["K"]), mask=g["f_mask"])
    inl = g["f_mask"].astype(bool)
    assert n == int(g["n_good"]) and np.abs(R - g["R"]).max() < 1e-9 and np.abs(t - g["t"]).max() < 1e-9
    assert np.array_equal(pm[inl], g["pose_mask"]) and pm[~inl].sum() == 0 and n == int(pm.sum())
    good = g["pose_mask"].astype(bool)
    rel = np.abs(X[inl][good] - g["X"][good]).max(1) / np.abs(g["X"][good]).max(1)
    assert rel.max() < 1e-5 and np.all(X[~pm.astype(bool)] == 0)
    assert abs(np.linalg.det(R) - 1.0) < 1e-12 and abs(np.linalg.norm(t) - 1.0) < 1e-12
    assert np.abs(R - g["R_true"]).max() < 1e-2 and np.abs(t - g["t_true"]).max() < 5e-2     # vs the true motion (cv2's F is a raw 7-point model)
    Ek = g["K"].T @ g["F"] @ g["K"]
    assert np.abs(E - Ek / np.linalg.norm(Ek)).max() < 1e-12


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pose_oracle_vs_cv2_live(seed):
    p1, p2, gt, _ = synth.two_view_correspondences(800, outlier_frac=0.3, seed=300 + seed)
    F, m, ninl, _ = ro.ransac_f(p1, p2, solver=7, seed=seed, lo=True)
    K = synth.K_INTR
    n, R, t, E, pm, X = ro.two_view_pose(p1, p2, F, _cam8(K), mask=m)
    inl = m.astype(bool)
    nc, Rc, tc, mc, Xc = cv2.recoverPose(K.T @ F @ K, p1[inl].astype(np.float64), p2[inl].astype(np.float64), K, distanceThresh=50.0)
    assert n == nc and np.abs(R - Rc).max() < 1e-9 and np.abs(t - tc.ravel()).max() < 1e-9
    assert np.array_equal(pm[inl], (mc.ravel() > 0).astype(np.uint8))
    # swapped images: the inverse motion
    n2, R2, t2, _, _, _ = ro.two_view_pose(p2, p1, F.T, _cam8(K), mask=m)
    assert n2 == n and np.abs(R2 - R.T).max() < 1e-6 and np.abs(t2 + R.T @ t).max() < 1e-6


def test_pose_oracle_degenerate():
    p1, p2, _, _ = synth.two_view_correspondences(50, outlier_frac=0.0, seed=1)
    n, R, t, E, pm, X = ro.two_view_pose(p1, p2, np.zeros((3, 3)), _cam8(synth.K_INTR))
    assert n == 0 and not R.any() and not t.any() and pm.sum() == 0
    n, R, t, E, pm, X = ro.two_view_pose(p1[:0], p2[:0], np.eye(3), _cam8(synth.K_INTR))
    assert n == 0


def test_lo_eigen_solver_agrees_with_numpy():
    """The LO refit's smallest-eigenvector solver (Cholesky + 16 inverse iterations, shared by the F and H oracles and
    restated value for value in csrc/ransac_common.cuh) against numpy.linalg.eigh; tolerance 1e-7 for eigenvalue gaps
    lambda_9 / lambda_8 <= 0.3 (the error contracts by that ratio per iteration), and exact rank deficiency is handled."""
    import ctypes as C

    L = ro.lib()
    L.sfm_oracle_smallest_eigvec9.restype = C.c_int
    rng = np.random.default_rng(0)
    for t in range(100):
        Q, _ = np.linalg.qr(rng.normal(size=(9, 9)))
        ev = np.sort(10.0 ** rng.uniform(-6, 3, 9))
        ev[0] = ev[1] * (rng.uniform(0.01, 0.3) if t % 3 else 0.0)
        A = np.ascontiguousarray(((Q * ev) @ Q.T + ((Q * ev) @ Q.T).T) / 2)
        x = np.zeros(9)
        assert L.sfm_oracle_smallest_eigvec9(A.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p)) == 1
        v = np.linalg.eigh(A)[1][:, 0]
        assert min(np.abs(x - v).max(), np.abs(x + v).max()) < 1e-7 and abs(np.linalg.norm(x) - 1) < 1e-12
    z = np.zeros((9, 9))
    assert L.sfm_oracle_smallest_eigvec9(z.ctypes.data_as(C.c_void_p), np.zeros(9).ctypes.data_as(C.c_void_p)) == 0


def test_triangulation_mirror_property_is_exact():
    """csrc/pose.cu triangulates once per rotation and derives the (R, -t) candidate as the exact negative: every IEEE
    operation of the DLT / Jacobi chain is sign-symmetric.  Pinned here on the oracle's triangulation (which the kernel
    equals bit for bit): X(R, -t) == -X(R, t) for random rotations, baselines and image points, to the last bit."""
    import ctypes as C

    L = ro.lib()
    L.sfm_oracle_triangulate.restype = C.c_int
    rng = np.random.default_rng(3)
    for _ in range(300):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        R = np.ascontiguousarray(q * np.sign(np.linalg.det(q)))
        t = rng.normal(size=3)
        t /= np.linalg.norm(t)
        xy = np.ascontiguousarray(rng.uniform(-0.6, 0.6, 4))
        Xa, Xb = np.zeros(3), np.zeros(3)
        args = lambda tt, X: (R.ctypes.data_as(C.c_void_p), np.ascontiguousarray(tt).ctypes.data_as(C.c_void_p),  # noqa: E731
                              xy.ctypes.data_as(C.c_void_p), C.c_double(1e30), X.ctypes.data_as(C.c_void_p))
        L.sfm_oracle_triangulate(*args(t, Xa))
        L.sfm_oracle_triangulate(*args(-t, Xb))
        assert np.array_equal(Xb, -Xa) and np.isfinite(Xa).all()


def test_pose_oracle_invariances():
    """Properties of the two-view step that do not depend on OpenCV: the recovered motion is invariant to the scale and
    sign of F and to a change of image resolution (pixels and intrinsics scaled together); triangulated points satisfy
    both projection equations."""
    p1, p2, gt, _ = synth.two_view_correspondences(600, outlier_frac=0.2, seed=404)
    F, m, ninl, _ = ro.ransac_f(p1, p2, solver=8, seed=3, lo=True)
    K = synth.K_INTR
    n0, R0, t0, E0, pm0, X0 = ro.two_view_pose(p1, p2, F, _cam8(K), mask=m)
    for scale in (-1.0, 3.5, -1e-3):
        n1, R1, t1, _, pm1, _ = ro.two_view_pose(p1, p2, scale * F, _cam8(K), mask=m)
        assert n1 == n0 and np.abs(R1 - R0).max() < 1e-9 and np.abs(t1 - t0).max() < 1e-9 and np.array_equal(pm1, pm0)
    s = 0.5                                                            # half-resolution images
    Ks = K.copy()
    Ks[:2] *= s
    Fs = np.diag([1 / s, 1 / s, 1.0]) @ F @ np.diag([1 / s, 1 / s, 1.0])
    n2, R2, t2, _, pm2, X2 = ro.two_view_pose(p1 * s, p2 * s, Fs, _cam8(Ks), mask=m)
    assert n2 == n0 and np.abs(R2 - R0).max() < 1e-6 and np.abs(t2 - t0).max() < 1e-6
    good = pm0.astype(bool) & gt
    x1 = (K @ X0[good].astype(np.float64).T).T
    x2 = (K @ (R0 @ X0[good].astype(np.float64).T + t0[:, None])).T
    assert np.median(np.abs(x1[:, :2] / x1[:, 2:3] - p1[good])) < 1.0 and np.median(np.abs(x2[:, :2] / x2[:, 2:3] - p2[good])) < 1.0
    assert abs(np.linalg.det(R0) - 1) < 1e-12 and np.abs(R0 @ R0.T - np.eye(3)).max() < 1e-12
    # E = [t]x R up to scale and sign
    tx = np.array([[0, -t0[2], t0[1]], [t0[2], 0, -t0[0]], [-t0[1], t0[0], 0]])
    Et = tx @ R0
    Et /= np.linalg.norm(Et)
    assert min(np.abs(Et - E0).max(), np.abs(Et + E0).max()) < 5e-2     # E0 comes from a noisy F: not exactly essential


# ------------------------------------------------------------------------------------ ORB descriptor stage
def _textured(rng, h, w):
    import cv2

    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(img.astype(int) + rng.normal(0, 10, (h, w)).astype(int), 0, 255).astype(np.uint8)


def test_orb_pattern_is_what_cv2_answers_to_probe_images(golden_dir):
    """The committed sampling pattern equals what the probe-image recovery measures from the installed cv2 right now
    (tools/recover_orb_pattern.py; the reference extracts with cv2.ORB, code/feature_matching.py:42-45)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("recover_orb_pattern", os.path.join(ROOT, "tools", "recover_orb_pattern.py"))
    rec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rec)
    pat = np.load(os.path.join(golden_dir, "orb_pattern.npz"))["pattern"]
    assert pat.shape == (256, 4) and pat.dtype == np.int8 and np.abs(pat).max() <= 15
    for axis in (0, 1):
        sc = rec.scan_axis(axis)
        for i in range(256):
            up, down = sc[False][i], sc[True][i]
            a, b = int(pat[i, axis]), int(pat[i, 2 + axis])
            if a < b:
                assert up == (a - 2, b + 3) and down is None
            elif a > b:
                assert down == (b - 2, a + 3) and up is None
            else:
                assert up is None and down is None
    inc = open(os.path.join(ROOT, "sfm-project_b200", "csrc", "orb_pattern.inc")).read()
    vals = [int(v) for v in re.findall(r"-?\d+", inc.split("\n", 1)[1])]
    assert vals == pat.reshape(-1).tolist()                              # the table compiled into the library is this pattern


def test_orb_oracle_stages_equal_cv2(golden_dir):
    """Every stage of the restated descriptor pipeline against the cv2 call it restates: INTER_LINEAR_EXACT resize through the
    product's 8.8 tables, the float32 blur against sepFilter2D, and whole descriptors against cv2.ORB on the reference-run golden
    images and on textured images of other sizes (all eight octaves)."""
    import cv2

    from oracle import orb_oracle
    from sfm_b200 import orb as porb

    rng = np.random.default_rng(5)
    # resize: tables of the product host code, applied with plain integer numpy
    for (h, w) in ((240, 320), (333, 517), (1080, 1920)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        sizes = porb.level_sizes(w, h)
        prev = img
        for k in range(1, 8):
            dw, dh = sizes[k]
            xt, yt = porb.resize_table(prev.shape[1], dw).astype(np.int64), porb.resize_table(prev.shape[0], dh).astype(np.int64)
            s = prev.astype(np.int64)
            x1, y1 = np.minimum(xt[:, 0] + 1, prev.shape[1] - 1), np.minimum(yt[:, 0] + 1, prev.shape[0] - 1)
            hl = (256 - xt[:, 1])[None] * s[:, xt[:, 0]] + xt[:, 1][None] * s[:, x1]
            v = (256 - yt[:, 1])[:, None] * hl[yt[:, 0]] + yt[:, 1][:, None] * hl[y1]
            mine = ((v + (1 << 15)) >> 16).astype(np.uint8)
            ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
            assert np.array_equal(mine, ref), (h, w, k)
            prev = ref
    # blur
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    for (h, w) in ((97, 131), (600, 800)):
        img = _textured(rng, h, w) if h > 100 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(orb_oracle.blur_level(img), cv2.sepFilter2D(img, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101))
    src = open(os.path.join(ROOT, "sfm-project_b200", "csrc", "orb.cu")).read()
    taps = [int(v, 16) for v in re.findall(r"(0x3[de][0-9a-f]{6})u", src)][:4]
    assert taps == [int(np.float32(v).view(np.uint32)) for v in k.ravel()[:4]]      # the kernel's taps are cv2's, bit for bit
    # descriptors
    pat = np.load(os.path.join(golden_dir, "orb_pattern.npz"))["pattern"]
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    for n, img in enumerate(g["images"]):
        kp = cv2.ORB_create().detect(img, None)
        assert np.array_equal(orb_oracle.describe(img, kp, pat), g[f"des{n}"])
    img = _textured(rng, 480, 640)
    kp, des = cv2.ORB_create(nfeatures=1500).detectAndCompute(img, None)
    assert sorted({k_.octave for k_ in kp}) == list(range(8)) and np.array_equal(orb_oracle.describe(img, kp, pat), des)


def test_orb_detection_oracle_equals_cv2(golden_dir):
    """The restated keypoint half of cv2.ORB (FAST-9/16 score + non-maximum suppression, border rule, retainBest through
    libstdc++'s nth_element / partition, Harris, intensity-centroid angle, level scaling) returns cv2.ORB_create().detect's
    keypoints -- position, size, angle, response, octave AND order -- on the reference-run golden images and on a textured image;
    each building block against the cv2 call it restates."""
    import cv2

    from oracle import orb_detect_oracle as od

    rng = np.random.default_rng(8)
    img = _textured(rng, 240, 320)
    xs, ys, resp = od.fast_keypoints(img)
    ref = cv2.FastFeatureDetector_create(20, True).detect(img, None)
    assert [(int(k.pt[0]), int(k.pt[1]), k.response) for k in ref] == list(zip(xs.tolist(), ys.tolist(), resp.tolist())) and len(ref) > 1000
    y, x = rng.normal(0, 1000, 5000).astype(np.float32), rng.normal(0, 1000, 5000).astype(np.float32)
    y[:40], x[40:80] = 0, 0
    assert np.array_equal(od.fast_atan2(y, x), np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32))
    assert od.features_per_level() == [109, 90, 75, 63, 52, 44, 36, 31]
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    for im in list(g["images"]) + [_textured(rng, 480, 640)]:
        kp = cv2.ORB_create().detect(im, None)
        ref = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kp], np.float32)
        assert np.array_equal(od.detect(im), ref)
