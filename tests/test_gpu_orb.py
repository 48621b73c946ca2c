"""GPU parity of ORB's descriptor stage (stage 1): pyramid, blur and rBRIEF tests for keypoints cv2
detected, bit for bit against cv2 itself -- the reference's extraction call is ``orb.detectAndCompute(gray, None)``
(code/feature_matching.py:42-45).  Every call goes through the C ABI (sfm_orb_resize / sfm_orb_blur / sfm_orb_describe)."""
import os
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)
pytestmark = pytest.mark.gpu

import cv2  # noqa: E402

import sfm_b200  # noqa: E402
from sfm_b200 import orb  # noqa: E402


def _textured(rng, h, w):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(img.astype(int) + rng.normal(0, 10, (h, w)).astype(int), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("hw", [(240, 320), (333, 517), (1080, 1920)])
def test_pyramid_and_blur_equal_cv2(hw):
    """Every raw level equals the cv2.resize(INTER_LINEAR_EXACT) chain and every blurred level equals
    cv2.sepFilter2D with the float32 Gaussian kernel (the call GaussianBlur makes for ORB's sub-matrix levels)."""
    h, w = hw
    rng = np.random.default_rng(h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8) if h < 300 else _textured(rng, h, w)
    d = orb.OrbDescriber(w, h)
    d.pyramid(img)
    torch.cuda.synchronize()
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    prev = img
    for lv, (lw, lh) in enumerate(d.sizes):
        if lv:
            prev = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(d.raw[lv].cpu().numpy(), prev), f"level {lv}"
        assert np.array_equal(d.blur[lv].cpu().numpy(), cv2.sepFilter2D(prev, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)), f"blur {lv}"


def test_descriptors_equal_cv2_on_the_golden_images(golden_dir):
    """Descriptors of every keypoint of the stored images == the reference run's (cv2.ORB_create().detectAndCompute)."""
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    for n, img in enumerate(g["images"]):
        kp = cv2.ORB_create().detect(img, None)
        des = orb.describe(img, kp).cpu().numpy()
        assert des.shape == g[f"des{n}"].shape and np.array_equal(des, g[f"des{n}"])


def test_descriptors_equal_cv2_live_all_octaves_and_straight_into_a_bank():
    rng = np.random.default_rng(11)
    timings = {}
    for (h, w, nf) in ((480, 640, 1500), (1080, 1920, 500), (1080, 1920, 5000)):
        img = _textured(rng, h, w)
        o = cv2.ORB_create(nfeatures=nf)
        t0 = time.perf_counter()
        kp, ref = o.detectAndCompute(img, None)
        t1 = time.perf_counter()
        kd = o.detect(img, None)
        t2 = time.perf_counter()
        des = orb.describe(img, kd)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        assert len(kd) == len(kp) and sorted({k.octave for k in kp}) == list(range(8))
        assert np.array_equal(des.cpu().numpy(), ref)
        timings[(h, w, nf)] = (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2))
    print("ms per image (cv2 detectAndCompute | cv2 detect | GPU pyramid + blur + descriptors incl. upload):", timings)
    # straight into the rows of a Hamming bank: matching two images from it equals the drop-in on cv2's own descriptors
    img1, img2 = _textured(rng, 480, 640), None
    img2 = np.roll(img1, 5, axis=1)
    o = cv2.ORB_create()
    (k1, d1), (k2, d2) = o.detectAndCompute(img1, None), o.detectAndCompute(img2, None)
    bank = sfm_b200.DescriptorBank(2, 512, metric="hamming")
    bank.put(0, orb.describe(img1, k1).unsqueeze(0))
    bank.put(1, orb.describe(img2, k2).unsqueeze(0))
    q, t, d = sfm_b200.match_pairs_hamming(bank, [[0, 1]], 26).to_host()[0]
    bf = sorted(cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(d1, d2), key=lambda m: m.distance)
    bf = [m for m in bf if m.distance < 26]
    assert [m.queryIdx for m in bf] == q.tolist() and [m.trainIdx for m in bf] == t.tolist() and len(q) > 20
    with pytest.raises(ValueError):
        orb.describe(img1, np.array([[5.0, 5.0, 0.0, 0.0]], np.float32))          # a hand-made keypoint on the border


def test_dropin_uses_gpu_descriptors_and_still_equals_the_reference(golden_dir, monkeypatch):
    """extract_and_match with the whole extraction on the GPU (the default), with cv2's detector + GPU descriptors, and with
    cv2's extraction (descriptors uploaded) returns the reference function's list every time."""
    import feature_matching as fm

    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    imgs = g["images"]
    for gpu, det in ((True, True), (True, False), (False, False)):
        monkeypatch.setattr(fm, "GPU_DESCRIPTORS", gpu)
        monkeypatch.setattr(fm, "GPU_DETECTION", det)
        fm._ORB_CACHE.clear()
        fm._SEEN.clear()
        fm._SLOTS = None
        for i, j in ((0, 1), (1, 2), (2, 0)):
            m = fm.extract_and_match(imgs[i], imgs[j])
            assert [x.queryIdx for x in m] == g[f"q_{i}_{j}"].tolist() and [x.trainIdx for x in m] == g[f"t_{i}_{j}"].tolist()
            assert [x.distance for x in m] == g[f"d_{i}_{j}"].tolist()
        kind = type(fm._ORB_CACHE[next(iter(fm._ORB_CACHE))][1]).__name__
        assert kind == ("Tensor" if gpu else "ndarray")


def _cv2_rows(kp):
    return np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kp], np.float32).reshape(-1, 6)


def test_detection_equals_cv2_keypoints_and_order(golden_dir):
    """Stage 2: FAST + non-maximum suppression + retainBest + Harris + orientation on the GPU return cv2.ORB_create().detect's
    keypoints -- position, size, angle, response, octave and ORDER -- and detect_and_compute returns detectAndCompute's
    descriptors, on the reference-run golden images and on textured images up to 1080p."""
    g = np.load(os.path.join(golden_dir, "ref_orb_hamming.npz"))
    rng = np.random.default_rng(21)
    timings = {}
    for n, img in enumerate(list(g["images"]) + [_textured(rng, 480, 640), _textured(rng, 1080, 1920), _textured(rng, 333, 517)]):
        o = cv2.ORB_create()
        t0 = time.perf_counter()
        kp, des = o.detectAndCompute(img, None)
        t1 = time.perf_counter()
        orb.detect_and_compute(img)                                    # first use of this size: buffers
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rows, d = orb.detect_and_compute(img)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        assert np.array_equal(rows, _cv2_rows(kp)), f"image {n}: keypoints differ"
        assert np.array_equal(d.cpu().numpy(), des), f"image {n}: descriptors differ"
        timings[img.shape] = (round(1e3 * (t1 - t0), 2), round(1e3 * (t3 - t2), 2))
        if n < 3:
            assert np.array_equal(des, g[f"des{n}"])
    print("ms per image (cv2 detectAndCompute | GPU detect_and_compute incl. upload and host selection):", timings)
    blank = np.zeros((120, 160), np.uint8)
    rows, d = orb.detect_and_compute(blank)
    assert rows.shape == (0, 6) and d.shape == (0, 32)


def test_fast_and_selection_building_blocks():
    """sfm_orb_fast_detect == cv2.FastFeatureDetector (positions, scores, row-major order) inside the border, and the host
    selection == the oracle's use of the same libstdc++ algorithms, ties included."""
    from oracle import orb_detect_oracle as od

    rng = np.random.default_rng(5)
    img = _textured(rng, 300, 420)
    e = orb.OrbExtractor(420, 300)
    e.pyramid(img)
    import ctypes as C

    from sfm_b200 import _lib

    _lib.check(_lib.lib().sfm_orb_fast_detect(_lib.ptr(e.raw[0]), 420, 300, 420, 20, 31, _lib.ptr(e.score[0]), _lib.ptr(e.row_count[0]),
                                              C.c_void_p(e.totals.data_ptr()), _lib.ptr(e.xy[0]), _lib.ptr(e.resp[0]), _lib.current_stream_ptr()), "fast")
    n = int(e.totals[0].item())
    ref = [k for k in cv2.FastFeatureDetector_create(20, True).detect(img, None) if 31 <= k.pt[0] < 420 - 31 and 31 <= k.pt[1] < 300 - 31]
    assert n == len(ref) and n > 500
    assert e.xy[0][:n].cpu().numpy().tolist() == [[int(k.pt[0]), int(k.pt[1])] for k in ref]
    assert e.resp[0][:n].cpu().numpy().tolist() == [k.response for k in ref]
    sc = e.score[0].cpu().numpy()
    assert np.array_equal(sc, od.fast_scores(img).astype(np.uint8))
    r = rng.integers(20, 60, 4000).astype(np.float32)                   # FAST-like responses: many ties at the threshold
    for npts in (0, 1, 100, 218, 4000, 5000):
        assert np.array_equal(orb.retain_best(r, npts), od.retain_best(r, npts))
