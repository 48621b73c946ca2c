"""GPU parity tests of the two stages downstream of RANSAC-F: batched RANSAC homography
(csrc/ransac_h.cu) and two-view pose recovery + triangulation (csrc/pose.cu).  Bit-exact against the seeded C oracles
(oracle/ransac_h.c, oracle/pose.c), which tests/test_oracle_pinned.py pins against cv2.findHomography / cv2.recoverPose;
plus the committed cv2 golden vectors and cv2 run live."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

import sfm_b200  # noqa: E402
from oracle import ransac_oracle as ro  # noqa: E402
from sfm_b200 import ransac as rs  # noqa: E402
from sfm_b200 import synth  # noqa: E402

cv2 = pytest.importorskip("cv2")

H_CASES = [(500, 0.3), (2000, 0.5), (3, 0.0), (1200, 0.6), (4, 0.0), (5, 0.0), (8192, 0.5)]


def _pack(pairs, cap=None):
    cap = cap or max(16, max(len(p1) for p1, _ in pairs))
    corr = np.zeros((len(pairs), cap, 4), np.float32)
    counts = np.zeros(len(pairs), np.int32)
    for k, (p1, p2) in enumerate(pairs):
        corr[k, : len(p1), :2], corr[k, : len(p1), 2:] = p1, p2
        counts[k] = len(p1)
    return torch.from_numpy(corr).cuda(), torch.from_numpy(counts)


@pytest.mark.parametrize("lo", [False, True])
def test_homography_bit_exact_vs_seeded_oracle(lo):
    data = [synth.planar_correspondences(n, outlier_frac=o, seed=60 + k) for k, (n, o) in enumerate(H_CASES)]
    data.append(synth.two_view_correspondences(1500, outlier_frac=0.2, seed=77))          # non-planar: weak H support
    corr, counts = _pack([(d[0], d[1]) for d in data])
    vb = rs.verify_h_corr(corr, counts, thr=3.0, confidence=0.995, max_iters=1024, lo=lo, seed=9)
    H, ninl, mask, iters = vb.F.cpu().numpy(), vb.n_inliers.cpu().numpy(), vb.mask.cpu().numpy(), vb.iters.cpu().numpy()
    for k, (p1, p2, gt, _) in enumerate(data):
        oH, om, on, oi = ro.ransac_h(p1, p2, pair_id=k, thr=3.0, max_iters=1024, confidence=0.995, seed=9, lo=lo)
        n = len(p1)
        assert ninl[k] == on and iters[k] == oi
        assert np.array_equal(mask[k, :n], om) and (mask[k, n:] == 0).all()
        if oH is None:
            assert ninl[k] == 0 and (H[k] == 0).all()
        else:
            assert np.array_equal(H[k], oH) and H[k][2, 2] == 1.0          # float64, bit for bit
    assert ninl[-1] < 0.5 * 1200


def test_homography_explicit_samples_packed_and_cv2():
    p1, p2, gt, Ht = synth.planar_correspondences(300, outlier_frac=0.2, seed=4)
    samples = np.random.default_rng(0).integers(0, 300, (256, 8)).astype(np.uint32)
    corr, counts = _pack([(p1, p2)])
    vb = rs.verify_h_corr(corr, counts, max_iters=256, confidence=1.0, samples=samples)
    oH, om, on, oi = ro.ransac_h(p1, p2, max_iters=256, confidence=1.0, samples=samples)
    assert int(vb.n_inliers[0]) == on and int(vb.iters[0]) == oi == 256
    assert np.array_equal(vb.mask[0, :300].cpu().numpy(), om) and np.array_equal(vb.F[0].cpu().numpy(), oH)
    # statistical against cv2 (own RNG) and the ground truth: IoU vs truth >= cv2's, transfer residual of true inliers
    ious, cious = [], []
    sets = [synth.planar_correspondences(2000, outlier_frac=0.5, seed=210 + s) for s in range(4)]
    corr, counts = _pack([(d[0], d[1]) for d in sets])
    vb = rs.verify_h_corr(corr, counts, thr=3.0, confidence=0.995, max_iters=2000, lo=True, seed=3)
    for k, (q1, q2, g, Htrue) in enumerate(sets):
        Hc, mc = cv2.findHomography(q1, q2, cv2.RANSAC, 3.0, maxIters=2000, confidence=0.995)
        m = vb.mask[k, :2000].cpu().numpy()
        ious.append(ro.iou(m, g))
        cious.append(ro.iou(mc.ravel(), g))
        Hk = vb.F[k].cpu().numpy()
        assert np.median(ro.transfer_err(Hk, q1[g], q2[g])) < 1.5          # px^2 at 0.5 px noise
        assert ro.iou(m, mc.ravel()) >= 0.98                                # north_star's IoU gate, applied to H
    assert np.mean(ious) >= np.mean(cious) - 0.005 and np.mean(ious) > 0.97
    # packed entry point == strided
    import ctypes as C
    from sfm_b200 import _lib
    tot = [len(d[0]) for d in sets]
    off = torch.tensor(np.concatenate([[0], np.cumsum(tot)]).astype(np.int32)).cuda()
    flat = torch.cat([corr[k, : tot[k]] for k in range(len(sets))]).contiguous()
    prm = rs.ransac_params(thr=3.0, confidence=0.995, max_iters=2000, solver="8pt", lo=True, seed=3)
    Hp = torch.zeros((len(sets), 9), dtype=torch.float64, device="cuda")
    nin = torch.zeros(len(sets), dtype=torch.int32, device="cuda")
    mk = torch.zeros(sum(tot), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().sfm_ransac_h_packed(_lib.ptr(flat), _lib.ptr(off), len(sets), 2048, None, None, None, C.byref(prm), _lib.ptr(Hp),
                                              _lib.ptr(nin), _lib.ptr(mk), None, _lib.current_stream_ptr()), "sfm_ransac_h_packed")
    assert torch.equal(Hp.view(-1, 3, 3), vb.F) and torch.equal(nin, vb.n_inliers)
    assert torch.equal(mk, torch.cat([vb.mask[k, : tot[k]] for k in range(len(sets))]))


def _cam8(K):
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]] * 2, np.float64)


def test_pose_bit_exact_vs_oracle_and_cv2():
    cases = [(800, 0.3), (2000, 0.5), (50, 0.0), (8192, 0.4), (12, 0.0)]
    data = [synth.two_view_correspondences(n, outlier_frac=o, seed=300 + k) for k, (n, o) in enumerate(cases)]
    corr, counts = _pack([(d[0], d[1]) for d in data])
    vb = sfm_b200.verify_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=1024, solver="7pt", lo=True, seed=5)
    K = synth.K_INTR
    cam = rs.camera_rows(K, None, len(data))
    assert np.array_equal(cam[0], _cam8(K))
    pb = rs.recover_pose_corr(corr, counts, vb.F, cam, mask=vb.mask)
    F, fm = vb.F.cpu().numpy(), vb.mask.cpu().numpy()
    R, t, E, ng = pb.R.cpu().numpy(), pb.t.cpu().numpy(), pb.E.cpu().numpy(), pb.n_good.cpu().numpy()
    pm, X = pb.mask.cpu().numpy(), pb.points.cpu().numpy()
    Ps = synth.make_cameras(2, baseline=6.0)
    Rt, tt = synth.relative_pose(Ps[0], Ps[1])
    for k, (p1, p2, gt, _) in enumerate(data):
        n = len(p1)
        on, oR, ot, oE, om, oX = ro.two_view_pose(p1, p2, F[k], cam[k], mask=fm[k, :n])
        assert ng[k] == on and np.array_equal(R[k], oR) and np.array_equal(t[k], ot) and np.array_equal(E[k], oE)
        assert np.array_equal(pm[k, :n], om) and (pm[k, n:] == 0).all()
        assert np.array_equal(X[k, :n], oX) and (X[k, n:] == 0).all()            # float32, bit for bit
        inl = fm[k, :n].astype(bool)
        nc, Rc, tc, mc, Xc = cv2.recoverPose(K.T @ F[k] @ K, p1[inl].astype(np.float64), p2[inl].astype(np.float64), K, distanceThresh=50.0)
        assert ng[k] == nc and np.abs(R[k] - Rc).max() < 1e-9 and np.abs(t[k] - tc.ravel()).max() < 1e-9
        assert np.array_equal(pm[k, :n][inl], (mc.ravel() > 0).astype(np.uint8))
        good = mc.ravel() > 0
        Xc = (Xc[:3] / Xc[3]).T
        assert (np.abs(X[k, :n][inl][good] - Xc[good]).max(1) / np.abs(Xc[good]).max(1)).max() < 1e-5
        if n >= 500:
            assert np.abs(R[k] - Rt).max() < 3e-2 and np.abs(t[k] - tt).max() < 8e-2      # the scene's true motion
        # triangulated points of true inliers reproject near their observation (each image allows 3 px to the epipolar line)
        sel = pm[k, :n].astype(bool) & gt
        x = (K @ X[k, :n][sel].astype(np.float64).T).T
        assert np.median(np.abs(x[:, :2] / x[:, 2:3] - p1[sel])) < 1.0 and np.abs(x[:, :2] / x[:, 2:3] - p1[sel]).max() < 8.0


def test_pose_golden_no_mask_no_model_and_packed(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_two_view.npz"))
    corr, counts = _pack([(g["pts1"], g["pts2"]), (g["pts1"], g["pts2"]), (g["pts1"][:0], g["pts2"][:0])], cap=512)
    F = torch.zeros((3, 3, 3), dtype=torch.float64)
    F[0] = torch.from_numpy(g["F"])                               # pair 1: no model, pair 2: no correspondences
    F[2] = torch.from_numpy(g["F"])
    mask = torch.zeros((3, 512), dtype=torch.uint8)
    mask[0, :500] = torch.from_numpy(g["f_mask"])
    mask[1, :500] = 1
    pb = rs.recover_pose_corr(corr, counts, F, rs.camera_rows(g["K"], g["K"], 3), mask=mask)
    inl = g["f_mask"].astype(bool)
    assert int(pb.n_good[0]) == int(g["n_good"])
    assert np.abs(pb.R[0].cpu().numpy() - g["R"]).max() < 1e-9 and np.abs(pb.t[0].cpu().numpy() - g["t"]).max() < 1e-9
    assert np.array_equal(pb.mask[0, :500].cpu().numpy()[inl], g["pose_mask"])
    for k in (1, 2):
        assert int(pb.n_good[k]) == 0 and not pb.R[k].any() and not pb.t[k].any() and not pb.mask[k].any() and not pb.points[k].any()
    # mask=None uses every correspondence (outliers vote too, the pose survives at 30 % outliers)
    pb2 = rs.recover_pose_corr(corr[:1], counts[:1], F[:1], rs.camera_rows(g["K"]), mask=None)
    assert np.abs(pb2.R[0].cpu().numpy() - g["R"]).max() < 1e-9 and int(pb2.n_good[0]) >= int(g["n_good"])
    with pytest.raises(ValueError):
        rs.recover_pose_corr(corr[:1], counts[:1], F[:1], rs.camera_rows(g["K"]), distance_thresh=0.0)
    with pytest.raises(ValueError):
        Ks = g["K"].copy(); Ks[0, 1] = 0.5
        rs.camera_rows(Ks)


def test_module_api_and_scene_graph_classification():
    """geometric_verification: find_homography / recover_pose / two_view_geometry keep cv2's return conventions, and the
    H-vs-F inlier ratio separates a planar pair from a general one."""
    import geometric_verification as gv

    K = synth.K_INTR
    g1, g2, ggt, _ = synth.two_view_correspondences(1500, outlier_frac=0.3, seed=11)
    q1, q2, qgt, Ht = synth.planar_correspondences(1500, outlier_frac=0.3, seed=12)
    H, hm = gv.find_homography(q1, q2, lo=True, seed=2)
    assert H.shape == (3, 3) and H.dtype == np.float64 and H[2, 2] == 1.0 and hm.shape == (1500, 1) and hm.dtype == np.uint8
    oH, om, on, _ = ro.ransac_h(q1, q2, lo=True, seed=2)
    assert np.array_equal(H, oH) and np.array_equal(hm.ravel(), om)
    assert gv.find_homography(q1[:3], q2[:3])[0] is None and gv.find_homographies([], []) == []
    tv = gv.two_view_geometry(g1, g2, K, seed=4)
    assert tv["config"] == gv.CALIBRATED and tv["n_inliers"] > 1000 and tv["n_inliers_h"] < 0.8 * tv["n_inliers"]
    assert tv["n_pose"] >= tv["n_inliers"] - 5 and abs(np.linalg.det(tv["R"]) - 1) < 1e-9 and tv["points3d"].shape == (1500, 3)
    on, oR, ot, _, opm, oX = ro.two_view_pose(g1, g2, tv["F"], _cam8(K), mask=tv["inlier_mask"])
    assert tv["n_pose"] == on and np.array_equal(tv["R"], oR) and np.array_equal(tv["t"], ot)
    assert np.array_equal(tv["in_front"].ravel(), opm) and np.array_equal(tv["points3d"], oX)
    depth = tv["points3d"][tv["in_front"].ravel().astype(bool) & ggt][:, 2]
    assert depth.min() > 0 and np.isfinite(depth).all()
    tp = gv.two_view_geometry(q1, q2, None, seed=4)
    assert tp["config"] == gv.PLANAR_OR_PANORAMIC and "R" not in tp
    assert gv.two_view_geometry(g1[:10], g2[:10], K)["config"] == gv.DEGENERATE
    cls = gv.classify_pairs([100, 100, 10], [50, 90, 10])
    assert list(cls) == [gv.UNCALIBRATED, gv.PLANAR_OR_PANORAMIC, gv.DEGENERATE]


def test_pipeline_with_homography_and_pose_stages():
    """match_and_verify(homography=True, intrinsics=K): the optional stages run on the same packed correspondences and
    agree bit for bit with the oracles fed the pipeline's own matches; the base results do not change."""
    from oracle import match_oracle as mo

    sc = synth.make_scene(4, 1024, seed=8)
    bank = sfm_b200.DescriptorBank(4, 1024)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = sfm_b200.exhaustive_pairs(4)
    K = synth.K_INTR
    base = sfm_b200.match_and_verify(bank, pairs, fetch=True, max_iters=256, seed=3, lo=True).to_host()
    for batch in (2048, 4):                                             # single batch, and two output sets alternating
        res = sfm_b200.match_and_verify(bank, pairs, fetch=True, max_iters=256, seed=3, lo=True, homography=True, intrinsics=K,
                                        pair_batch=batch)
        h = res.to_host()
        for k in ("matches", "inlier", "F", "n_inliers", "offsets"):
            assert np.array_equal(h[k], base[k])
        assert torch.equal(res.H.cpu(), torch.from_numpy(h["H"])) and torch.equal(res.n_pose.cpu(), torch.from_numpy(h["n_pose"]))
        for p, (i, j) in enumerate(pairs):
            sl = slice(h["offsets"][p], h["offsets"][p + 1])
            q, t = h["matches"][sl, 0], h["matches"][sl, 1]
            p1, p2 = sc.xy[i][q], sc.xy[j][t]
            oH, ohm, ohn, _ = ro.ransac_h(p1, p2, pair_id=p, max_iters=256, confidence=0.99, seed=3, lo=True,
                                          stop_target=rs.h_stop_target(int(h["n_inliers"][p]), 0.8))
            assert h["n_inliers_h"][p] == ohn and np.array_equal(h["inlier_h"][sl], ohm) and np.array_equal(h["H"][p], oH)
            on, oR, ot, _, opm, oX = ro.two_view_pose(p1, p2, h["F"][p], _cam8(K), mask=h["inlier"][sl])
            assert h["n_pose"][p] == on and np.array_equal(h["R"][p], oR) and np.array_equal(h["t"][p], ot)
            assert np.array_equal(h["in_front"][sl], opm) and np.array_equal(h["points3d"][sl], oX)
            Rt, tt = synth.relative_pose(sc.P[i], sc.P[j])
            assert np.abs(h["R"][p] - Rt).max() < 3e-2
        summ = res.to_host(with_matches=False)
        assert "points3d" not in summ and np.array_equal(summ["R"], h["R"])
    import geometric_verification as gv
    assert all(c == gv.CALIBRATED for c in gv.classify_pairs(h["n_inliers"], h["n_inliers_h"], calibrated=True))


def test_homography_stop_target_bit_exact_and_cheap_on_general_pairs():
    """stop_target (the scene-graph question "can H explain 80 % of what F explains?"): a general pair stops after the first
    32 hypotheses instead of the whole budget, a planar pair is unaffected; bit-exact against the oracle with the same target."""
    g1, g2, _, _ = synth.two_view_correspondences(2000, outlier_frac=0.2, seed=91)
    q1, q2, _, _ = synth.planar_correspondences(2000, outlier_frac=0.2, seed=92)
    corr, counts = _pack([(g1, g2), (q1, q2)])
    vf = sfm_b200.verify_corr(corr, counts, max_iters=512, solver="7pt", lo=True, seed=2)
    tgt = (vf.n_inliers.to(torch.float32) * 0.8).to(torch.int32)
    plain = rs.verify_h_corr(corr, counts, max_iters=2000, confidence=0.99, seed=4)
    fast = rs.verify_h_corr(corr, counts, max_iters=2000, confidence=0.99, seed=4, stop_target=tgt)
    assert int(plain.iters[0]) == 2000 and int(fast.iters[0]) == 32              # general pair: H cannot reach the target
    assert int(fast.iters[1]) == int(plain.iters[1]) and torch.equal(fast.F[1], plain.F[1])     # planar pair: same run
    for k, (p1, p2) in enumerate(((g1, g2), (q1, q2))):
        oH, om, on, oi = ro.ransac_h(p1, p2, pair_id=k, max_iters=2000, confidence=0.99, seed=4, stop_target=int(tgt[k]))
        assert int(fast.iters[k]) == oi and int(fast.n_inliers[k]) == on and np.array_equal(fast.mask[k, :2000].cpu().numpy(), om)
        assert np.array_equal(fast.F[k].cpu().numpy(), oH)
    import geometric_verification as gv
    cls = gv.classify_pairs(vf.n_inliers.cpu().numpy(), fast.n_inliers.cpu().numpy())
    assert list(cls) == [gv.UNCALIBRATED, gv.PLANAR_OR_PANORAMIC]
    with pytest.raises(ValueError):
        rs.verify_h_corr(corr, counts, stop_target=tgt[:1])


def test_optional_stages_edge_cases():
    """Pairs with no or too few matches, per-image intrinsics, summaries without fetch: nothing crashes, empty results are
    all-zero, and per-image camera rows reach the kernel in pair order."""
    sc = synth.make_scene(4, 512, seed=17)
    sc.desc[3] = synth.sift_like(np.random.default_rng(9), 512)          # image 3 shares nothing: (almost) no matches with it
    bank = sfm_b200.DescriptorBank(4, 512)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = sfm_b200.exhaustive_pairs(4)
    K = synth.K_INTR
    Ks = np.stack([K, K, K, K])
    res = sfm_b200.match_and_verify(bank, pairs, max_iters=128, seed=1, lo=True, homography=True, intrinsics=Ks, min_inliers=8)
    s = res.to_host(with_matches=False)
    weak = np.array([3 in p for p in pairs.tolist()])
    assert (s["n_matches"][weak] < 8).all() and (s["n_matches"][~weak] > 100).all()
    assert (s["n_inliers"][weak] == 0).all() and (s["n_inliers_h"][weak] == 0).all() and (s["n_pose"][weak] == 0).all()
    assert not s["R"][weak].any() and not s["t"][weak].any() and not s["H"][weak].any() and not s["F"][weak].any()
    assert (s["n_pose"][~weak] > 0.9 * s["n_inliers"][~weak]).all()
    assert np.allclose(np.linalg.det(s["R"][~weak]), 1.0, atol=1e-9) and np.allclose(np.linalg.norm(s["t"][~weak], axis=1), 1.0, atol=1e-12)
    # the same intrinsics as [n,4] rows and as one shared matrix give identical results
    rows = np.array([[K[0, 0], K[1, 1], K[0, 2], K[1, 2]]] * 4)
    for intr in (rows, K):
        s2 = sfm_b200.match_and_verify(bank, pairs, max_iters=128, seed=1, lo=True, homography=True, intrinsics=intr, min_inliers=8).to_host(with_matches=False)
        assert np.array_equal(s2["R"], s["R"]) and np.array_equal(s2["n_pose"], s["n_pose"]) and np.array_equal(s2["H"], s["H"])
    # different cameras per image: the per-pair rows follow the pair list (doubling image 1's focal length changes only its pairs)
    Kd = Ks.copy()
    Kd[1, 0, 0] *= 2.0
    Kd[1, 1, 1] *= 2.0
    s3 = sfm_b200.match_and_verify(bank, pairs, max_iters=128, seed=1, lo=True, intrinsics=Kd, min_inliers=8).to_host(with_matches=False)
    touched = np.array([1 in p for p in pairs.tolist()])
    assert np.array_equal(s3["R"][~touched], s["R"][~touched]) and not np.array_equal(s3["R"][touched & ~weak], s["R"][touched & ~weak])
    with pytest.raises(ValueError):
        sfm_b200.match_and_verify(bank, pairs, intrinsics=np.zeros((2, 3, 3)))
    import geometric_verification as gv
    assert gv.recover_poses([], [], [], K) == [] and gv.find_homographies([], []) == []
    n, R, t, m, X = gv.recover_pose(None, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), K)
    assert n == 0 and not R.any() and m.shape == (0, 1) and X.shape == (0, 3)
