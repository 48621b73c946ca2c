"""GPU parity tests of batched RANSAC-F (K4): bit-exact against the seeded C oracle (same samples), statistical
against cv2 (IoU vs ground truth >= cv2's), and the cv2 conventions of the returned F / mask."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

import sfm_b200  # noqa: E402
from oracle import cv2_ref  # noqa: E402
from oracle import ransac_oracle as ro  # noqa: E402
from sfm_b200 import synth  # noqa: E402

CASES = [(500, 0.3), (2000, 0.5), (6, 0.0), (1200, 0.6), (7, 0.0), (8, 0.0), (8192, 0.5)]


def _batch(cases, seed0=40, cap=None):
    data, cs, counts = [], [], []
    cap = cap or max(16, max(n for n, _ in cases))
    for k, (n, outl) in enumerate(cases):
        p1, p2, gt, F = synth.two_view_correspondences(n, outlier_frac=outl, seed=seed0 + k)
        data.append((p1, p2, gt, F))
        c = np.zeros((cap, 4), np.float32)
        c[:n, :2], c[:n, 2:] = p1, p2
        cs.append(c)
        counts.append(n)
    return data, torch.from_numpy(np.stack(cs)).cuda(), torch.tensor(counts, dtype=torch.int32)


@pytest.mark.parametrize("solver", ["7pt", "8pt"])
@pytest.mark.parametrize("lo", [False, True])
@pytest.mark.parametrize("score", ["sym_epipolar", "sampson"])
def test_ransac_bit_exact_vs_seeded_oracle(solver, lo, score):
    data, corr, counts = _batch(CASES)
    vb = sfm_b200.verify_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=1024, solver=solver, score=score, lo=lo, seed=9)
    F, ninl = vb.F.cpu().numpy(), vb.n_inliers.cpu().numpy()
    mask, iters = vb.mask.cpu().numpy(), vb.iters.cpu().numpy()
    for k, (p1, p2, gt, _) in enumerate(data):
        oF, om, on, oi = ro.ransac_f(p1, p2, pair_id=k, solver=int(solver[0]), score=0 if score == "sym_epipolar" else 1,
                                     thr=3.0, max_iters=1024, confidence=0.99, seed=9, lo=lo)
        n = len(p1)
        assert ninl[k] == on and iters[k] == oi
        assert np.array_equal(mask[k, :n], om) and (mask[k, n:] == 0).all()
        if oF is None:
            assert ninl[k] == 0 and (F[k] == 0).all()
        else:
            assert np.array_equal(F[k], oF)                  # float64, bit for bit
            assert F[k][2, 2] == 1.0


def test_ransac_explicit_samples_bit_exact():
    data, corr, counts = _batch([(300, 0.2)], seed0=4)
    samples = np.random.default_rng(0).integers(0, 300, (256, 8)).astype(np.uint32)
    vb = sfm_b200.verify_corr(corr, counts, max_iters=256, confidence=1.0, solver="8pt", samples=samples)
    p1, p2, _, _ = data[0]
    oF, om, on, oi = ro.ransac_f(p1, p2, solver=8, max_iters=256, confidence=1.0, samples=samples)
    assert int(vb.n_inliers[0]) == on and int(vb.iters[0]) == oi == 256
    assert np.array_equal(vb.mask[0, :300].cpu().numpy(), om) and np.array_equal(vb.F[0].cpu().numpy(), oF)


def test_ransac_statistical_vs_cv2_and_ground_truth():
    """IoU against synthetic ground truth is at least cv2's; Sampson residual of the returned F on true inliers is
    small (tolerance 1e-4 px^2 is for noise-free data: checked in the next test)."""
    ious, cious = [], []
    for seed in range(4):
        p1, p2, gt, _ = synth.two_view_correspondences(2000, outlier_frac=0.5, seed=200 + seed)
        import geometric_verification as gv

        F, mask = gv.verify_pair(p1, p2, thr=3.0, confidence=0.99, max_iters=2000, solver="7pt", lo=True, seed=seed)
        Fc, mc = cv2_ref.find_fundamental(p1, p2, 3.0, 0.99, 2000)
        assert F is not None and F.shape == (3, 3) and mask.shape == (2000, 1) and mask.dtype == np.uint8
        assert abs(F[2, 2] - 1.0) < 1e-12
        assert np.median(ro.sampson_err(F, p1[gt], p2[gt])) < 1.0
        ious.append(ro.iou(mask, gt))
        cious.append(ro.iou(mc, gt))
        print(f"seed {seed}: IoU vs gt {ious[-1]:.4f} (cv2 {cious[-1]:.4f}), IoU vs cv2 {ro.iou(mask, mc):.4f}")
    assert np.mean(ious) >= np.mean(cious) - 0.005 and np.mean(ious) > 0.97


def test_ransac_noise_free_residual_tolerance():
    p1, p2, gt, Ft = synth.two_view_correspondences(500, outlier_frac=0.3, seed=77, pixel_sigma=0.0)
    import geometric_verification as gv

    F, mask = gv.verify_pair(p1, p2, thr=1.0, max_iters=2000, solver="8pt", lo=True)
    assert ro.iou(mask, gt) >= 0.98
    assert ro.sampson_err(F, p1[gt], p2[gt]).max() < 1e-4          # px^2, the tolerance north_star states


def test_ransac_degenerate_and_api_conventions():
    import cv2
    import geometric_verification as gv

    p1, p2, _, _ = synth.two_view_correspondences(6, outlier_frac=0.0, seed=1)
    F, mask = gv.verify_pair(p1, p2)
    assert F is None and mask.shape == (6, 1) and mask.sum() == 0           # cv2 returns (None, None) here
    F, mask = gv.verify_pair(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert F is None and mask.shape == (0, 1)
    same = np.tile(np.array([[10.0, 20.0]], np.float32), (50, 1))
    F, mask = gv.verify_pair(same, same, max_iters=256)
    assert F is None and mask.sum() == 0
    with pytest.raises(ValueError):
        gv.verify_pair(np.zeros((5, 3), np.float32), np.zeros((5, 3), np.float32))
    with pytest.raises(ValueError):
        gv.verify_pair(p1, p2, solver="5pt")
    p1, p2, gt, _ = synth.two_view_correspondences(400, outlier_frac=0.2, seed=3)
    # accepts [M,1,2] and float64 like cv2 does
    Fa, ma = gv.verify_pair(p1.reshape(-1, 1, 2), p2.astype(np.float64), seed=2)
    Fb, mb = gv.verify_pair(p1, p2, seed=2)
    assert np.array_equal(Fa, Fb) and np.array_equal(ma, mb)
    kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p1]
    kp2 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p2]
    ms = [cv2.DMatch(i, i, 0.0) for i in range(400)]
    Fm, inl = gv.verify_matches(kp1, kp2, ms, seed=2)
    assert np.array_equal(Fm, Fb) and len(inl) == int(mb.sum())
    assert gv.verify_matches(kp1, kp2, []) == (None, [])


def test_match_and_verify_end_to_end_vs_oracles():
    """The whole hot path on a small exhaustive run: every pair's matches, F, mask equal the oracles'."""
    from oracle import match_oracle as mo

    sc = synth.make_scene(4, 1024, seed=5)
    bank = sfm_b200.DescriptorBank(4, 1024)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = synth.exhaustive_pairs(4)
    res = sfm_b200.match_and_verify(bank, pairs, max_iters=512, seed=3, lo=True, pair_batch=4, fetch=True)   # also exercises batching
    h = res.to_host()
    assert h["pairs"].tolist() == pairs.tolist()
    for p, (i, j) in enumerate(pairs):
        q, t, d = mo.match_l2(sc.desc[i], sc.desc[j], ratio=0.75)
        sl = slice(h["offsets"][p], h["offsets"][p + 1])
        assert np.array_equal(h["matches"][sl, 0], q) and np.array_equal(h["matches"][sl, 1], t)
        F, mask, ninl, iters = ro.ransac_f(sc.xy[i][q], sc.xy[j][t], pair_id=p, solver=7, max_iters=512, seed=3, lo=True)
        assert h["n_inliers"][p] == ninl and np.array_equal(h["inlier"][sl], mask) and np.array_equal(h["F"][p], F)
        Ft = sc.true_fundamental(i, j)
        gt = sc.point[i][q] == sc.point[j][t]
        assert np.median(ro.sym_epipolar_err(Ft, sc.xy[i][q][mask.astype(bool)], sc.xy[j][t][mask.astype(bool)])) < 2.0
        assert ro.iou(mask, gt) > 0.95


def test_plan_path_prefilter_fetch_modes_and_mutual():
    """The execution plan (packed buffers, side-stream copies) returns the same result whatever the batch size,
    with or without the sweep's prefilter, as copies or as pinned views; the resident summaries agree with the
    fetched ones; mutual matching equals the oracle's cross-checked ratio matches."""
    from oracle import match_oracle as mo

    sc = synth.make_scene(5, 2048, seed=8)
    bank = sfm_b200.DescriptorBank(5, 2048)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = synth.exhaustive_pairs(5)
    kw = dict(max_iters=256, seed=9, solver="8pt")
    a = sfm_b200.match_and_verify(bank, pairs, fetch=True, prefilter=True, pair_batch=3, **kw).to_host()
    b = sfm_b200.match_and_verify(bank, pairs, fetch=True, prefilter=False, pair_batch=16, **kw).to_host()
    v = sfm_b200.match_and_verify(bank, pairs, fetch="view", pair_batch=16, **kw)
    c = {k: np.array(x) for k, x in v.to_host().items()}
    assert v.d2h_bytes == 4 * 11 + 13 * len(c["matches"]) + 10 * 80
    r = sfm_b200.match_and_verify(bank, pairs, pair_batch=4, **kw)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    rs = r.to_host(with_matches=False)
    for k in ("n_matches", "n_inliers", "F", "iters"):
        assert np.array_equal(rs[k], a[k]), k
    with pytest.raises(ValueError):
        r.to_host()
    assert a["n_matches"].min() > 50 and (a["n_inliers"] > 0).all()
    m = sfm_b200.match_and_verify(bank, pairs[:3], fetch=True, mutual=True, **kw).to_host()
    for p, (i, j) in enumerate(pairs[:3]):
        q, t, d = mo.match_l2(sc.desc[i], sc.desc[j], ratio=0.75, mutual=True)
        sl = slice(m["offsets"][p], m["offsets"][p + 1])
        assert np.array_equal(m["matches"][sl, 0], q) and np.array_equal(m["matches"][sl, 1], t) and np.array_equal(m["matches"][sl, 2], d)
        F, mask, ninl, iters = ro.ransac_f(sc.xy[i][q], sc.xy[j][t], pair_id=p, solver=8, max_iters=256, seed=9)
        assert m["n_inliers"][p] == ninl and np.array_equal(m["inlier"][sl], mask) and np.array_equal(m["F"][p], F)
    e = sfm_b200.match_and_verify(bank, np.zeros((0, 2), np.int32), fetch=True).to_host()
    assert len(e["matches"]) == 0 and e["offsets"].tolist() == [0]


def test_ransac_packed_equals_strided_and_wide_pairs():
    """sfm_ransac_f_packed on back-to-back correspondences == the strided call; a pair wider than the 4096 points
    kept in shared memory (tail read through L2) is bit-exact against the oracle too."""
    import ctypes as C

    from sfm_b200 import _lib
    from sfm_b200.ransac import ransac_params

    cases = [(500, 0.3), (6, 0.0), (6000, 0.5), (0, 0.0), (1200, 0.6)]
    data, corr, counts = _batch([c for c in cases if c[0] > 0] + [(8, 0.0)], cap=8192)
    counts = counts.clone()
    counts[-1] = 0                                              # an empty pair in the middle of the packed list
    order = [0, 1, 2, 4, 3]
    corr, counts = corr[order].contiguous(), counts[order].contiguous()
    vb = sfm_b200.verify_corr(corr, counts, solver="7pt", max_iters=384, seed=4, lo=True)
    n = counts.numpy()
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
    packed = torch.cat([corr[k, : n[k]] for k in range(len(n))]).contiguous()
    P, total = len(n), int(off[-1])
    F = torch.zeros((P, 9), dtype=torch.float64, device="cuda")
    ninl = torch.zeros(P, dtype=torch.int32, device="cuda")
    iters = torch.zeros(P, dtype=torch.int32, device="cuda")
    mask = torch.full((total + 16,), 9, dtype=torch.uint8, device="cuda")
    prm = ransac_params(solver="7pt", max_iters=384, seed=4, lo=True)
    off_d = torch.from_numpy(off).cuda()
    _lib.check(_lib.lib().sfm_ransac_f_packed(_lib.ptr(packed), _lib.ptr(off_d), P, 8192, None, None, C.byref(prm), _lib.ptr(F),
                                              _lib.ptr(ninl), _lib.ptr(mask), _lib.ptr(iters), _lib.current_stream_ptr()), "packed")
    assert torch.equal(F.view(P, 3, 3), vb.F) and torch.equal(ninl, vb.n_inliers) and torch.equal(iters, vb.iters)
    mask = mask.cpu().numpy()
    assert (mask[total:] == 9).all()
    for k in range(P):
        assert np.array_equal(mask[off[k]: off[k + 1]], vb.mask[k, : n[k]].cpu().numpy())
    # the wide pair (6000 > 4096 points in shared memory) against the C oracle
    p1, p2, gt, _ = data[2]
    oF, om, on, oi = ro.ransac_f(p1, p2, pair_id=2, solver=7, max_iters=384, seed=4, lo=True)
    assert int(vb.n_inliers[2]) == on and np.array_equal(vb.mask[2, :6000].cpu().numpy(), om) and np.array_equal(vb.F[2].cpu().numpy(), oF)


def test_streamed_host_job_equals_resident_run():
    """match_and_verify_host (chunked upload on a side stream, pairs re-ordered by the chunk of max(i, j)) gives every
    pair the result of match_and_verify on a resident bank."""
    sc = synth.make_scene(7, 1536, seed=14)
    pairs = synth.exhaustive_pairs(7)
    bank = sfm_b200.DescriptorBank(7, 1536)
    bank.put(0, sc.desc, xy=sc.xy)
    kw = dict(max_iters=256, seed=2, solver="7pt", lo=True)
    a = sfm_b200.match_and_verify(bank, pairs, fetch=True, **kw).to_host()
    desc_pin, xy_pin = torch.from_numpy(sc.desc).pin_memory(), torch.from_numpy(sc.xy).pin_memory()
    for n_chunks, batch in ((3, 2048), (7, 4), (1, 2048)):
        bank2 = sfm_b200.DescriptorBank(7, 1536)
        res, order = sfm_b200.match_and_verify_host(desc_pin, xy_pin, pairs, bank=bank2, n_chunks=n_chunks, pair_batch=batch, fetch=True, **kw)
        b = res.to_host()
        assert sorted(order.tolist()) == list(range(len(pairs))) and np.array_equal(b["pairs"], pairs[order])
        for k, p in enumerate(order):
            sa = slice(a["offsets"][p], a["offsets"][p + 1])
            sb = slice(b["offsets"][k], b["offsets"][k + 1])
            assert np.array_equal(a["matches"][sa], b["matches"][sb]) and np.array_equal(a["inlier"][sa], b["inlier"][sb])
            assert np.array_equal(a["F"][p], b["F"][k]) and a["n_inliers"][p] == b["n_inliers"][k] and a["iters"][p] == b["iters"][k]


def test_reference_pair_list_from_batched_result():
    """to_reference_pairs rebuilds the `pair_matches` list of code/pipeline.py:36-49 (Pair.img_inx_1 / img_inx_2 / matches as
    list[cv2.DMatch]) from the batched result, over the reference's ORDERED pair enumeration."""
    from oracle import match_oracle as mo

    sc = synth.make_scene(3, 1024, seed=21)
    sc.desc[2] = synth.sift_like(np.random.default_rng(5), 1024)          # image 2 shares nothing: its pairs have (almost) no matches
    bank = sfm_b200.DescriptorBank(3, 1024)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = sfm_b200.ordered_pairs(3)
    assert pairs.tolist() == [[0, 1], [0, 2], [1, 0], [1, 2], [2, 0], [2, 1]]
    h = sfm_b200.match_and_verify(bank, pairs, fetch=True, max_iters=128).to_host()
    plist = sfm_b200.to_reference_pairs(h)
    assert all(type(p.matches[0]).__name__ == "DMatch" for p in plist)
    for p in plist:
        q, t, d = mo.match_l2(sc.desc[p.img_inx_1], sc.desc[p.img_inx_2], ratio=0.75)
        assert [m.queryIdx for m in p.matches] == q.tolist() and [m.trainIdx for m in p.matches] == t.tolist()
        assert [m.distance for m in p.matches] == [float(x) for x in np.sqrt(d.astype(np.float32))]
    kept = {(p.img_inx_1, p.img_inx_2) for p in plist}
    assert {(0, 1), (1, 0)} <= kept
    strong = sfm_b200.to_reference_pairs(h, inliers_only=True, min_matches=50, as_dmatch=False)
    assert {(p.img_inx_1, p.img_inx_2) for p in strong} == {(0, 1), (1, 0)}
    assert all(len(p.matches) >= 50 and p.matches.shape[1] == 3 for p in strong)


def test_sustained_multi_batch_runs_stay_identical():
    """Stress for the sweep's handshake protocol: many back-to-back batches, uploads on a side stream, repeated jobs on one
    bank.  (A three-issuer variant of the sweep passed every single-batch test and failed exactly here: consecutive uses
    of a TMEM accumulator were issued by different threads and a parity wait aliased one mbarrier phase back.)  Every
    repetition must reproduce the single-batch result bit for bit."""
    sc = synth.make_scene(50, 8192, seed=33, desc_sigma=6.0)
    pairs = synth.exhaustive_pairs(50)                                   # the bench's job: 1,225 pairs x 64 train tiles
    bank = sfm_b200.DescriptorBank(50, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    kw = dict(max_iters=64, seed=2, solver="8pt")
    ref = sfm_b200.match_and_verify(bank, pairs, fetch=True, **kw).to_host()
    desc_pin, xy_pin = torch.from_numpy(sc.desc).pin_memory(), torch.from_numpy(sc.xy).pin_memory()
    for rep in range(6):
        n_chunks, batch = ((3, 512), (4, 256), (2, 640))[rep % 3]
        res, order = sfm_b200.match_and_verify_host(desc_pin, xy_pin, pairs, bank=bank, n_chunks=n_chunks, pair_batch=batch, fetch="view", **kw)
        b = res.to_host()
        assert np.array_equal(b["n_matches"], ref["n_matches"][order]) and np.array_equal(b["n_inliers"], ref["n_inliers"][order])
        assert np.array_equal(b["F"], ref["F"][order])
        k = int(np.argmax(order == 100))
        assert np.array_equal(b["matches"][b["offsets"][k]: b["offsets"][k + 1]], ref["matches"][ref["offsets"][100]: ref["offsets"][101]])
    again = sfm_b200.match_and_verify(bank, pairs, pair_batch=301, fetch=True, **kw).to_host()
    assert np.array_equal(again["matches"], ref["matches"]) and np.array_equal(again["inlier"], ref["inlier"])


def test_host_job_keeps_callers_ransac_streams():
    """match_and_verify_host(pair_ids=...): the RANSAC sample stream of the caller's pair k is pair_ids[k] whatever order
    the pairs are processed in, so a sharded / chunked run reproduces match_and_verify(pair_ids=...) on a resident bank."""
    sc = synth.make_scene(6, 1024, seed=19)
    pairs = synth.exhaustive_pairs(6)
    ids = np.arange(len(pairs)) * 7 + 1000
    bank = sfm_b200.DescriptorBank(6, 1024)
    bank.put(0, sc.desc, xy=sc.xy)
    kw = dict(max_iters=128, seed=4, solver="8pt")
    a = sfm_b200.match_and_verify(bank, pairs, pair_ids=ids, fetch=True, **kw).to_host()
    plain = sfm_b200.match_and_verify(bank, pairs, fetch=True, **kw).to_host()
    assert not np.array_equal(a["F"], plain["F"])                         # different streams -> different raw models
    res, order = sfm_b200.match_and_verify_host(torch.from_numpy(sc.desc).pin_memory(), torch.from_numpy(sc.xy).pin_memory(), pairs,
                                                n_chunks=3, pair_batch=4, pair_ids=ids, fetch=True, **kw)
    b = res.to_host()
    assert np.array_equal(b["F"], a["F"][order]) and np.array_equal(b["n_inliers"], a["n_inliers"][order])
    assert np.array_equal(b["iters"], a["iters"][order])


def test_row_sink_path_on_one_gpu_equals_fetch():
    """The device-side gather of the sharded run (``dist.match_and_verify_sharded``: rows pushed into a region batch by batch,
    three output sets, a push lag of two batches) with a world of ONE process: the region is a local array and everything in
    it must equal the host-fetch result -- for one, two and many batches, and when the region is reused by the next job.
    (The two-GPU form of this test is tests/test_gpu_multi.py; the driver's one-GPU run covers the same code through this one.)"""
    from sfm_b200 import dist as sdist

    sc = synth.make_scene(6, 2048, seed=5)
    pairs = synth.exhaustive_pairs(6)                                    # 15 pairs
    bank = sfm_b200.DescriptorBank(6, 2048)
    bank.put(0, sc.desc, xy=sc.xy)
    prm = dict(ratio=0.75, max_iters=256, solver="8pt", seed=4, lo=True)
    base = sfm_b200.match_and_verify(bank, pairs, fetch=True, **prm).to_host()
    for batch in (64, 8, 2, 3, 2):
        out, local = sdist.match_and_verify_sharded(bank, pairs, pair_batch=batch, **prm)
        torch.cuda.synchronize()
        for k in ("n_matches", "n_inliers", "iters", "F"):
            assert np.array_equal(out[k].cpu().numpy(), base[k]), (batch, k)
        start = out["row_start"].cpu().numpy()
        m, inl = out["matches"].cpu().numpy(), out["inlier"].cpu().numpy()
        for p in range(len(pairs)):
            a, b = base["offsets"][p], base["offsets"][p + 1]
            assert np.array_equal(m[start[p]: start[p] + (b - a)], base["matches"][a:b]), (batch, p)
            assert np.array_equal(inl[start[p]: start[p] + (b - a)], base["inlier"][a:b]), (batch, p)
    for reg in bank.__dict__.get("_gather_regions", {}).values():
        reg.close()
