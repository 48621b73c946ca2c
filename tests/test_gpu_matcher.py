"""GPU parity tests of the matcher (K1/K2/K3/K5) against the CPU oracle and the golden fixtures.
Everything goes through the C ABI (lib/libsfm_b200.so).  Bit-exact: integer work."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # collected on CPU boxes, skipped there
    pytest.skip("needs a GPU", allow_module_level=True)

import sfm_b200  # noqa: E402
from oracle import match_oracle as mo  # noqa: E402
from sfm_b200 import matcher, synth  # noqa: E402


def _pair(n1, n2, seed, planted=0.5):
    rng = np.random.default_rng(seed)
    B = synth.sift_like(rng, n2)
    A = synth.sift_like(rng, n1)
    k = int(min(n1, n2) * planted)
    if k:
        A[:k] = synth.observe(rng, B[rng.permutation(n2)[:k]])
    return A, B


def _check_knn(knn, A, B):
    i1, d1, i2, d2 = mo.l2_knn2(A, B)
    g = knn[: len(A)]
    assert np.array_equal(g[:, 0], i1) and np.array_equal(g[:, 1], d1)
    assert np.array_equal(g[:, 2], i2) and np.array_equal(g[:, 3], d2)
    assert (knn[len(A):] == -1).all()


def test_bank_pack_matches_restatement():
    A, B = _pair(700, 900, 0)
    bank = sfm_b200.build_bank([A, B])
    assert bank.feat_stride == 1024 and bank.counts.cpu().tolist()[:2] == [700, 900]
    nb = ((B.astype(np.int64) - 128) ** 2).sum(1)
    assert np.array_equal(bank.norms[1, :900].cpu().numpy(), nb)
    assert (bank.norms[1, 900:].cpu().numpy() == 0).all()
    ext = bank.section("ext").cpu().numpy()
    t0 = ext[(bank.feat_stride // 128) * 4096:][:4096].reshape(2, 128, 16)
    e = np.concatenate([t0[0], t0[1]], axis=1).astype(np.int64)
    w = np.array([255] * 24 + [1] + [0] * 7)
    assert np.array_equal((e * w).sum(1), (1 << 20) - (nb[:128] >> 1))       # K-extension encodes H0 - floor(|b|^2/2)
    desc = bank.section("desc").cpu().numpy()[: 2 * 1024 * 128].reshape(2, 1024, 128)
    assert np.array_equal(desc[0, :700] ^ 0x80, A)


@pytest.mark.parametrize("mode,name", [(1, "main"), (2, "ext"), (0, "both")])
def test_tcgen05_tile_accumulators(mode, name):
    """Raw TMEM accumulators of one 256 x 128 tile: descriptor MMAs, K-extension MMA, and both."""
    A, B = _pair(700, 900, 1)
    bank = sfm_b200.build_bank([A, B])
    acc, _ = matcher.debug_tc_tile(bank, [0, 1], mode)
    a = A[:256].astype(np.int64) - 128
    b = B[:128].astype(np.int64) - 128
    dot = a @ b.T
    ext = (1 << 20) - ((b * b).sum(1) >> 1)
    want = {0: dot + ext[None, :], 1: dot, 2: np.broadcast_to(ext[None, :], dot.shape)}[mode]
    assert np.array_equal(acc.cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize("n1,n2,seed", [(700, 900, 0), (2048, 2048, 1), (300, 5000, 3), (1, 2, 4), (5, 1, 5), (257, 129, 6), (129, 33, 7)])
@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_knn2_vs_oracle(n1, n2, seed, impl):
    A, B = _pair(n1, n2, seed)
    bank = sfm_b200.build_bank([A, B])
    knn = sfm_b200.knn2(bank, [[0, 1], [1, 0], [0, 0]], impl=impl).cpu().numpy()
    _check_knn(knn[0], A, B)
    _check_knn(knn[1], B, A)
    _check_knn(knn[2], A, A)          # self match: D1 = 0 at the own index, duplicates tie-break low


def test_knn2_golden_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_l2_knn.npz"))
    bank = sfm_b200.build_bank([g["A"], g["B"]])
    for impl in ("tcgen05", "simt"):
        knn = sfm_b200.knn2(bank, [[0, 1]], impl=impl).cpu().numpy()[0, : len(g["A"])]
        assert knn[:, 0].tolist() == g["idx1"].tolist() and knn[:, 2].tolist() == g["idx2"].tolist()
        assert np.sqrt(knn[:, 1].astype(np.float32)).tolist() == g["d1"].tolist()       # bitwise == cv2's float32
        assert np.sqrt(knn[:, 3].astype(np.float32)).tolist() == g["d2"].tolist()
    mb = sfm_b200.match_pairs(bank, [[0, 1]], ratio=0.75, ratio_mode="cv2_f32")
    q, t, d = mb.to_host()[0]
    assert q.tolist() == g["ratio_q"].tolist() and t.tolist() == g["ratio_t"].tolist()
    assert np.sqrt(d.astype(np.float32)).tolist() == g["ratio_d"].tolist()
    q, t, d = sfm_b200.match_pairs(bank, [[0, 1]], ratio=None, mutual=True).to_host()[0]
    assert q.tolist() == g["cross_q"].tolist() and t.tolist() == g["cross_t"].tolist()
    # exact-integer ratio mode drops exactly the two crafted boundary rows (16*D1 == 9*D2)
    qi = sfm_b200.match_pairs(bank, [[0, 1]], ratio=0.75, ratio_mode="exact_int").to_host()[0][0]
    assert sorted(set(g["ratio_q"].tolist()) - set(qi.tolist())) == [202, 203]


def test_adversarial_ties_and_extremes():
    """Identical descriptors (every chunk maximum ties -> brute-force path), all-0 / all-255 rows (extreme norms)."""
    rng = np.random.default_rng(5)
    B = synth.sift_like(rng, 600)
    B[100:400] = B[7]                         # 300 identical train rows spread over several tiles
    B[500] = 0
    B[501] = 255
    A = synth.sift_like(rng, 300)
    A[0] = B[7]
    A[1] = 0
    A[2] = 255
    A[3:40] = B[7]
    bank = sfm_b200.build_bank([A, B])
    for impl in ("tcgen05", "simt"):
        knn = sfm_b200.knn2(bank, [[0, 1], [1, 0]], impl=impl).cpu().numpy()
        _check_knn(knn[0], A, B)
        _check_knn(knn[1], B, A)
    assert knn[0, 0].tolist() == [7, 0, 100, 0]


@pytest.mark.parametrize("mode,mutual", [("cv2_f32", False), ("cv2_f32", True), ("exact_int", False), (None, True)])
def test_filter_vs_oracle(mode, mutual):
    A, B = _pair(1500, 1800, 9)
    bank = sfm_b200.build_bank([A, B], keypoint_xy=[np.random.default_rng(1).uniform(0, 1000, (1500, 2)),
                                                    np.random.default_rng(2).uniform(0, 1000, (1800, 2))])
    mb = sfm_b200.match_pairs(bank, [[0, 1]], ratio=0.75 if mode else None, ratio_mode=mode, mutual=mutual)
    q, t, d = mb.to_host()[0]
    oq, ot, od = mo.match_l2(A, B, ratio=0.75 if mode else None, ratio_mode=mode or "cv2_f32", mutual=mutual)
    assert np.array_equal(q, oq) and np.array_equal(t, ot) and np.array_equal(d, od)
    corr = mb.corr[0, : len(q)].cpu().numpy()
    xy = bank.xy.cpu().numpy()
    assert np.array_equal(corr[:, :2], xy[0][q]) and np.array_equal(corr[:, 2:], xy[1][t])


def test_full_size_pair_properties():
    """BASELINE size (8192 x 8192): tcgen05 == SIMT bit for bit, planted matches are recovered, and the
    result is invariant under a permutation of the train rows (size-independent properties)."""
    sc = synth.make_scene(3, 8192, seed=11)
    bank = sfm_b200.DescriptorBank(3, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = [[0, 1], [1, 2], [2, 0]]
    kt = sfm_b200.knn2(bank, pairs, impl="tcgen05")
    ks = sfm_b200.knn2(bank, pairs, impl="simt")
    assert torch.equal(kt, ks)
    mb = sfm_b200.match_pairs(bank, pairs)
    q, t, d = mb.to_host()[0]
    same = sc.point[0][q] == sc.point[1][t]
    assert same.mean() > 0.99 and len(q) > 1500
    perm = np.random.default_rng(0).permutation(8192)
    desc2 = sc.desc.copy()
    desc2[1] = sc.desc[1][perm]
    bank2 = sfm_b200.DescriptorBank(3, 8192)
    bank2.put(0, desc2)
    k2 = sfm_b200.knn2(bank2, [[0, 1]]).cpu().numpy()[0]
    k1 = kt[0].cpu().numpy()
    assert np.array_equal(k1[:, 1], k2[:, 1]) and np.array_equal(k1[:, 3], k2[:, 3])          # distances unchanged
    untied = k1[:, 1] != k1[:, 3]
    assert np.array_equal(perm[k2[untied, 0]], k1[untied, 0])                                   # same neighbour


def test_hamming_vs_oracle_and_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    n = g["images"].shape[0]
    bank = sfm_b200.build_bank([g[f"des{k}"] for k in range(n)], metric="hamming")
    pairs = synth.ordered_pairs(n)                      # the reference's loop order, code/pipeline.py:38-40
    res = sfm_b200.match_pairs_hamming(bank, pairs, 26).to_host()
    for (i, j), (q, t, d) in zip(pairs, res):
        assert q.tolist() == g[f"q_{i}_{j}"].tolist() and t.tolist() == g[f"t_{i}_{j}"].tolist()
        assert d.astype(np.float32).tolist() == g[f"d_{i}_{j}"].tolist()
    rng = np.random.default_rng(3)
    b = rng.integers(0, 256, (1300, 32), dtype=np.uint8)
    a = rng.integers(0, 256, (777, 32), dtype=np.uint8)
    a[:300] = b[rng.permutation(1300)[:300]] ^ (rng.random((300, 32)) < 0.03).astype(np.uint8)
    a[7] = a[0]
    b[1299] = b[0]
    hb = sfm_b200.build_bank([a, b], metric="hamming")
    for p, (X, Y) in enumerate([(a, b), (b, a)]):
        q, t, d = sfm_b200.match_pairs_hamming(hb, [[p, 1 - p]], 40).to_host()[0]
        oq, ot, od = mo.match_hamming_reference(X, Y, 40)
        assert np.array_equal(q, oq) and np.array_equal(t, ot) and np.array_equal(d, od)


def test_dropin_extract_and_match_equals_reference_golden(golden_dir):
    """The drop-in `extract_and_match(gray_i, gray_j)` (code/feature_matching.py:41) on the stored images returns
    the same list[cv2.DMatch] the reference function produced in the build container."""
    import feature_matching as fm

    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    imgs = g["images"]
    for i in range(len(imgs)):
        for j in range(len(imgs)):
            if i == j:
                continue
            m = fm.extract_and_match(imgs[i], imgs[j])
            assert isinstance(m, list) and m and type(m[0]).__name__ == "DMatch"
            assert [x.queryIdx for x in m] == g[f"q_{i}_{j}"].tolist()
            assert [x.trainIdx for x in m] == g[f"t_{i}_{j}"].tolist()
            assert [x.distance for x in m] == g[f"d_{i}_{j}"].tolist()
            assert all(x.imgIdx == 0 for x in m)
    blank = np.zeros((120, 160), np.uint8)
    assert fm.extract_and_match(blank, imgs[0]) == [] and fm.extract_and_match(imgs[0], blank) == []
    d1 = fm.match_descriptors_l2(np.zeros((0, 128), np.uint8), np.zeros((4, 128), np.uint8))
    assert d1 == []


def test_unmodified_pipeline_loop_runs_on_dropin(golden_dir):
    """code/pipeline.py:36-47 restated verbatim around the drop-in module (the file itself cannot travel)."""
    import feature_matching as fm

    images = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))["images"]

    class Pair:
        img_inx_1 = -1
        img_inx_2 = -1
        matches = []

    pair_matches = []
    for i in range(images.shape[0]):
        for j in range(images.shape[0]):
            if i != j:
                match = fm.extract_and_match(images[i], images[j])
                if match:
                    pair = Pair()
                    pair.img_inx_1, pair.img_inx_2, pair.matches = i, j, match
                    pair_matches.append(pair)
    assert len(pair_matches) == 6 and pair_matches[0].img_inx_2 == 1


def test_errors_are_loud():
    A, B = _pair(64, 64, 0)
    bank = sfm_b200.build_bank([A, B])
    with pytest.raises(ValueError):
        sfm_b200.match_pairs(bank, [[0, 2]])
    with pytest.raises(ValueError):
        sfm_b200.match_pairs_hamming(bank, [[0, 1]])
    with pytest.raises(ValueError):
        bank.put(0, np.zeros((1, 64, 32), np.uint8))
    with pytest.raises(ValueError):
        sfm_b200.match_pairs(bank, [[0, 1]], ratio_mode="nope")
    assert sfm_b200.match_pairs(bank, np.zeros((0, 2), np.int32)).counts.numel() == 0


def test_filter_packed_equals_strided():
    """sfm_filter_matches_packed (count -> scan -> write) returns the strided filter's rows back to back, for ragged
    pairs including an empty one."""
    import ctypes as C

    from sfm_b200 import _lib

    rng = np.random.default_rng(21)
    descs = [synth.sift_like(rng, n) for n in (900, 1, 1300, 513)]
    descs[2][:400] = synth.observe(rng, descs[0][:400])
    descs[3][:300] = synth.observe(rng, descs[0][500:800])
    xy = [rng.uniform(0, 1000, (len(d), 2)) for d in descs]
    bank = sfm_b200.build_bank(descs, keypoint_xy=xy)
    pairs = [[0, 2], [1, 0], [2, 0], [0, 3], [3, 1], [2, 3]]
    for mutual in (False, True):
        mb = sfm_b200.match_pairs(bank, pairs, ratio=0.8, mutual=mutual)
        ref = mb.to_host()
        pairs_t = torch.tensor(pairs, dtype=torch.int32, device="cuda")
        P, cap = len(pairs), bank.feat_stride
        fwd = sfm_b200.knn2(bank, pairs_t)
        rev = sfm_b200.knn2(bank, pairs_t.flip(1).contiguous()) if mutual else None
        cnt = torch.zeros(P, dtype=torch.int32, device="cuda")
        off = torch.zeros(P + 1, dtype=torch.int32, device="cuda")
        m = torch.full((P * cap, 3), -7, dtype=torch.int32, device="cuda")
        c = torch.zeros((P * cap, 4), dtype=torch.float32, device="cuda")
        prm = matcher.filter_params(0.8, "cv2_f32", mutual)
        _lib.check(_lib.lib().sfm_filter_matches_packed(bank.handle, _lib.ptr(pairs_t), P, _lib.ptr(fwd), _lib.ptr(rev), C.byref(prm),
                                                        _lib.ptr(cnt), _lib.ptr(off), _lib.ptr(m), _lib.ptr(c),
                                                        _lib.current_stream_ptr()), "packed")
        cnt, off, m, c = cnt.cpu().numpy(), off.cpu().numpy(), m.cpu().numpy(), c.cpu().numpy()
        assert off[0] == 0 and np.array_equal(np.diff(off), cnt) and cnt.tolist() == [len(r[0]) for r in ref]
        assert (m[off[-1]:] == -7).all()                                      # nothing written past the packed end
        for p in range(P):
            rows = m[off[p]: off[p + 1]]
            assert np.array_equal(rows[:, 0], ref[p][0]) and np.array_equal(rows[:, 1], ref[p][1]) and np.array_equal(rows[:, 2], ref[p][2])
            assert np.array_equal(c[off[p]: off[p + 1]], mb.corr[p, : cnt[p]].cpu().numpy())


@pytest.mark.parametrize("mode,ratio", [("cv2_f32", 0.75), ("cv2_f32", 0.9), ("exact_int", 0.75)])
def test_prefilter_only_drops_rows_that_fail_the_ratio_test(mode, ratio):
    """The sweep's prefilter (bounds on D1, D2 from the tile maxima) may only blank rows that fail the exact ratio
    test; every other row is the exact kNN.  So filtering the prefiltered table gives the oracle's matches."""
    import ctypes as C

    from sfm_b200 import _lib

    A, B = _pair(3000, 2600, 31, planted=0.4)
    A[5] = B[7]                                  # exact duplicate: D1 = 0
    B[9] = B[8]                                  # duplicate train rows: D1 == D2 for their observers
    bank = sfm_b200.build_bank([A, B])
    pairs_t = torch.tensor([[0, 1], [1, 0]], dtype=torch.int32, device="cuda")
    exact = sfm_b200.knn2(bank, pairs_t).cpu().numpy()
    fp = matcher.filter_params(ratio, mode, False)
    prm = _lib.MatchParams()
    prm.prefilter_mode, prm.prefilter_ratio = fp.ratio_mode, fp.ratio
    prm.prefilter_num, prm.prefilter_den = int(fp.ratio_num), int(fp.ratio_den)
    out = torch.empty_like(torch.from_numpy(exact)).cuda()
    _lib.check(_lib.lib().sfm_match_knn2(bank.handle, _lib.ptr(pairs_t), 2, C.byref(prm), _lib.ptr(out),
                                         _lib.current_stream_ptr()), "knn2 prefilter")
    pre = out.cpu().numpy()
    for p, (X, Y) in enumerate([(A, B), (B, A)]):
        n = len(X)
        blank = pre[p, :n, 0] < 0
        keep = mo.ratio_keep(exact[p, :n, 1], exact[p, :n, 3], ratio, mode)
        assert not (blank & keep).any()                                      # never drops a row that passes
        assert np.array_equal(pre[p, :n][~blank], exact[p, :n][~blank])       # the rest is the exact table
        assert blank.sum() > 0.3 * n                                          # and it does drop most failing rows
        assert (pre[p, n:] == -1).all()


def test_high_res_32768_features_pair():
    """BASELINE configs[3] feature count (32,768 per image, 256 train tiles): tcgen05 == SIMT bit for bit on a ragged
    pair, and the prefiltered match list equals the plain one."""
    rng = np.random.default_rng(77)
    B = synth.sift_like(rng, 32768)
    A = synth.sift_like(rng, 30001)
    A[:9000] = synth.observe(rng, B[rng.permutation(32768)[:9000]])
    bank = sfm_b200.build_bank([A, B])
    assert bank.feat_stride == 32768
    kt = sfm_b200.knn2(bank, [[0, 1], [1, 0]], impl="tcgen05")
    ks = sfm_b200.knn2(bank, [[0, 1], [1, 0]], impl="simt")
    assert torch.equal(kt, ks)
    q, t, d = sfm_b200.match_pairs(bank, [[0, 1]]).to_host()[0]
    assert len(q) > 8000
    h = sfm_b200.match_and_verify(bank, [[0, 1]], fetch=True, max_iters=64).to_host()
    assert np.array_equal(h["matches"][:, 0], q) and np.array_equal(h["matches"][:, 1], t) and np.array_equal(h["matches"][:, 2], d)
    # ... and pinned to cv2 run live at this shape (about 5 s of CPU): the kNN table of the forward direction and the ratio matches
    from oracle import cv2_ref

    i1, d1, i2, d2 = cv2_ref.l2_knn2(A, B)
    k = kt[0, : len(A)].cpu().numpy()
    assert np.array_equal(k[:, 0], i1) and np.array_equal(k[:, 2], i2)
    assert np.array_equal(np.sqrt(k[:, 1].astype(np.float32)), d1) and np.array_equal(np.sqrt(k[:, 3].astype(np.float32)), d2)
    cq, ct, cd = cv2_ref.l2_ratio_match(A, B, 0.75)
    assert np.array_equal(q, cq) and np.array_equal(t, ct) and np.array_equal(np.sqrt(d.astype(np.float32)), cd)


def test_config0_two_view_equals_cv2_live():
    """BASELINE configs[0] shape (one two-view pair, 8192 SIFT-like descriptors per image): the GPU match list equals
    cv2.BFMatcher(NORM_L2).knnMatch(k=2) + `m.distance < 0.75 * n.distance` run live, element for element, including the
    float32 distances the drop-in reports."""
    from oracle import cv2_ref

    sc = synth.make_scene(2, 8192, seed=12)
    q, t, dist = cv2_ref.l2_ratio_match(sc.desc[0], sc.desc[1], 0.75)
    bank = sfm_b200.DescriptorBank(2, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    h = sfm_b200.match_and_verify(bank, [[0, 1]], ratio=0.75, fetch=True).to_host()
    assert np.array_equal(h["matches"][:, 0], q) and np.array_equal(h["matches"][:, 1], t)
    assert np.array_equal(np.sqrt(h["matches"][:, 2].astype(np.float32)), dist)
    import feature_matching as fm

    dm = fm.match_descriptors_l2(sc.desc[0], sc.desc[1], ratio=0.75)
    assert [m.queryIdx for m in dm] == q.tolist() and [m.trainIdx for m in dm] == t.tolist()
    assert [m.distance for m in dm] == [float(x) for x in dist]
    # verification of that pair: the inlier set agrees with the ground truth at least as well as cv2's own
    gt = sc.point[0][q] == sc.point[1][t]
    Fc, mc = cv2_ref.find_fundamental(sc.xy[0][q], sc.xy[1][t], 3.0, 0.99, 2000)
    from oracle import ransac_oracle as ro

    assert ro.iou(h["inlier"], gt) >= ro.iou(mc, gt) - 0.01


def test_extract_and_match_draw_equals_extract_and_match(monkeypatch, golden_dir):
    """code/feature_matching.py:15-37 on the GPU path with pyplot stubbed: the draw twin returns the same list as
    extract_and_match (and as the reference function's golden output)."""
    import cv2
    import feature_matching as fm

    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    shown = []

    class Plt:
        def imshow(self, img):
            shown.append(img.shape)

        def show(self):
            shown.append("show")

    monkeypatch.setattr(fm, "plt", Plt())
    img1, img2 = g["images"][0], g["images"][1]
    a = fm.extract_and_match(img1, img2)
    b = fm.extract_and_match_draw(img1, img2)
    assert [(m.queryIdx, m.trainIdx, m.distance) for m in a] == [(m.queryIdx, m.trainIdx, m.distance) for m in b] and len(a) > 0
    assert [m.queryIdx for m in b] == g["q_0_1"].tolist() and [m.distance for m in b] == g["d_0_1"].tolist()
    assert shown[-1] == "show" and len(shown[0]) == 3 and shown[0][1] == img1.shape[1] + img2.shape[1]


@pytest.mark.parametrize("mode,ratio,mutual,prefilter,max_d", [("cv2_f32", 0.75, False, True, 0), ("cv2_f32", 0.9, True, True, 0),
                                                                 ("exact_int", 0.75, False, True, 0), ("cv2_f32", 0.75, False, False, 90000),
                                                                 ("none", None, True, True, 0), ("exact_int", 0.6, True, False, 0)])
def test_fused_refinement_filter_equals_the_knn_table_path(mode, ratio, mutual, prefilter, max_d):
    """sfm_match_pairs_packed (refinement + ratio / mutual filter in one pass, no kNN table) returns exactly the arrays of
    sfm_match_knn2 + sfm_filter_matches_packed -- counts, offsets, (queryIdx, trainIdx, D) rows in order and the gathered
    keypoint coordinates -- on ragged images with planted matches, exact duplicates and an empty image; and the kNN-table path
    is the one pinned to the numpy oracle and to cv2 above."""
    rng = np.random.default_rng(4)
    base = synth.sift_like(rng, 4000)
    imgs, xys = [], []
    for n in (3000, 2600, 515, 1, 2048):
        d = synth.sift_like(rng, n)
        k = min(n, 1500)
        d[:k] = synth.observe(rng, base[rng.permutation(4000)[:k]])
        imgs.append(d)
        xys.append(rng.random((n, 2)).astype(np.float32) * 1000)
    imgs[1][9] = imgs[1][8]
    imgs[0][5] = imgs[1][7]
    bank = sfm_b200.build_bank(imgs + [None], keypoint_xy=xys + [None])
    pairs = [[0, 1], [1, 0], [2, 4], [4, 3], [3, 2], [0, 5], [5, 1], [1, 4]]
    got = matcher.match_pairs_packed(bank, pairs, ratio=ratio, ratio_mode=mode, mutual=mutual, prefilter=prefilter, fused=True, max_distance_sq=max_d)
    ref = matcher.match_pairs_packed(bank, pairs, ratio=ratio, ratio_mode=mode, mutual=mutual, prefilter=prefilter, fused=False, max_distance_sq=max_d)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    assert int(ref[0].sum()) > 1000 and int(ref[0][5]) == 0 and int(ref[0][6]) == 0
    assert len(matcher.match_pairs_packed(bank, np.zeros((0, 2), np.int32))[2]) == 0


@pytest.mark.parametrize("mode,ratio", [("cv2_f32", 0.75), ("cv2_f32", 0.999), ("exact_int", 0.98), ("exact_int", 0.5), ("none", None)])
def test_fused_early_stop_on_near_ties(mode, ratio):
    """The fused refinement stops after the record's first tile when the second-largest tile maximum proves that no other tile
    can change the nearest neighbour or the outcome of the ratio test.  Stress for exactly that decision: descriptors from a
    two-letter alphabet (distances collide all over the image, nearest and second nearest sit in different tiles with equal
    or adjacent distances, odd and even norms mix) and ratios next to 1, where one unit of the second distance flips the test.
    The fused form must return what the kNN-table form (pinned to the oracle and to cv2) returns."""
    rng = np.random.default_rng(11)
    imgs = []
    for n in (1500, 1300, 700):
        d = np.where(rng.random((n, 128)) < 0.2, 65, 0).astype(np.uint8)            # |b - 128|^2 odd or even depending on the row
        d[:: 7] = np.where(rng.random((len(d[:: 7]), 128)) < 0.2, 64, 1).astype(np.uint8)
        imgs.append(d)
    imgs[1][:400] = imgs[0][rng.permutation(1500)[:400]]                              # exact copies: distance 0 against near-copies
    flip = rng.integers(0, 128, 400)
    imgs[1][np.arange(400), flip] ^= 1                                                # ... and copies one unit away
    bank = sfm_b200.build_bank(imgs)
    pairs = [[0, 1], [1, 0], [0, 2], [2, 1], [1, 2]]
    for prefilter in (True, False):
        got = matcher.match_pairs_packed(bank, pairs, ratio=ratio, ratio_mode=mode, prefilter=prefilter, fused=True)
        ref = matcher.match_pairs_packed(bank, pairs, ratio=ratio, ratio_mode=mode, prefilter=prefilter, fused=False)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
    if mode != "none":
        assert 0 < int(ref[0].sum()) < sum(len(imgs[p[0]]) for p in pairs)
