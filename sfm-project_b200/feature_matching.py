"""Drop-in replacement for the reference's ``code/feature_matching.py``.

Put this directory on ``sys.path`` ahead of the reference's ``code/`` and the unmodified
``code/pipeline.py`` runs on the B200 path: it star-imports this module
(``from feature_matching import*``, code/pipeline.py:2) and calls ``extract_and_match(gray_i, gray_j)``
(code/pipeline.py:41).  Names, argument meaning and return conventions are the reference's
(code/feature_matching.py:9, :15, :41); the module also keeps ``os``, ``cv2``, ``np``, ``math``, ``plt``,
``bisect`` as public globals because pipeline.py uses ``os``/``cv2``/``np`` without importing them
(code/pipeline.py:14-19) and defines no ``__all__`` for the same reason.

What changed underneath:
* ``orb.detectAndCompute`` runs on the GPU, both halves bit-exact including the ORDER of the keypoints (csrc/orb.cu):
  FAST + non-maximum suppression + Harris + orientation, then pyramid blur and the 256 rotated intensity tests
  (``SFM_ORB_DETECTION=cv2`` keeps cv2's detector, ``SFM_ORB_DESCRIPTORS=cv2`` the whole of cv2's extraction).  Extraction is cached per image content, so the
  reference's N(N-1) pair loop extracts each image once instead of 2(N-1) times.
* ``cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match`` + ``sorted`` + ``distance < 26``
  (code/feature_matching.py:48-58) run on the GPU (csrc/hamming.cu) and return the identical
  ``list[cv2.DMatch]``.
There is no CPU fallback: without the CUDA library or a B200 the calls raise.
"""
import os
import cv2
import numpy as np
import math
import bisect
import hashlib
import weakref

try:  # matplotlib is optional (absent in this image); the reference imports it at module level
    import matplotlib.pyplot as plt
except Exception:  # pragma: no cover - depends on the environment
    class _NoPyplot:
        def __getattr__(self, name):
            raise ImportError("matplotlib is not installed; extract_and_match_draw cannot display")

    plt = _NoPyplot()

import sfm_b200 as _sfm

MAX_HAMMING_DISTANCE = 26          # code/feature_matching.py:29 and :55
GPU_DESCRIPTORS = os.environ.get("SFM_ORB_DESCRIPTORS", "gpu").lower() != "cv2"
GPU_DETECTION = GPU_DESCRIPTORS and os.environ.get("SFM_ORB_DETECTION", "gpu").lower() != "cv2"
_ORB_CACHE = {}
_ORB_CACHE_MAX = 4096


def read_img(path):
    # read image in grayscale (the 0 flag indicates grayscale) -- code/feature_matching.py:9-11
    return cv2.imread(path, 0)


def _check_image(gray, name):
    if not isinstance(gray, np.ndarray) or gray.ndim != 2 or gray.dtype != np.uint8:
        raise ValueError(f"{name} must be a 2-D uint8 grayscale image")


_SEEN = {}                         # id(array) -> (weakref, data pointer, shape, strides, sparse fingerprint, content key)


def _content_key(gray):
    """Cache key of an image.  The reference's loop passes the SAME ndarray objects again and again (rows of the array built
    at code/pipeline.py:19), so an object seen before -- same id, still alive, same buffer, same sparse fingerprint (~1 K
    pixels) -- reuses its key; only a new object pays the full-content hash (1-2 ms per 1080p image, far more than the GPU
    match itself)."""
    h, w = gray.shape
    probe = gray[:: max(1, h // 32), :: max(1, w // 32)].tobytes()
    ent = _SEEN.get(id(gray))
    if ent is not None and ent[0]() is gray and ent[1] == gray.ctypes.data and ent[2] == gray.shape and ent[3] == gray.strides \
            and ent[4] == probe:
        return ent[5]
    key = (gray.shape, hashlib.blake2b(np.ascontiguousarray(gray), digest_size=16).digest())
    try:
        if len(_SEEN) >= _ORB_CACHE_MAX:
            _SEEN.clear()
        _SEEN[id(gray)] = (weakref.ref(gray), gray.ctypes.data, gray.shape, gray.strides, probe, key)
    except TypeError:              # (an ndarray subclass without weak references: always hashed)
        pass
    return key


def _extract(gray):
    """cv2.ORB_create().detectAndCompute(gray, None) (code/feature_matching.py:42-45), cached per image."""
    key = _content_key(gray)
    hit = _ORB_CACHE.get(key)
    if hit is None:
        orb = cv2.ORB_create()
        if GPU_DETECTION and GPU_DESCRIPTORS:
            # detection AND descriptors on the device, the keypoints (and their order) identical to cv2's; rows (pt.x, pt.y, size, angle,
            # response, octave) -- cv2.KeyPoint objects are only built when something draws them
            kp, des = _sfm.orb.detect_and_compute(gray)
            kp = _KeypointRows(kp)
            des = des if len(kp) else None
        elif GPU_DESCRIPTORS:
            # orb.detect returns the keypoints of detectAndCompute; their descriptors are computed on the device and stay there
            kp = orb.detect(gray, None)
            des = _sfm.orb.describe(gray, kp) if len(kp) else None
        else:
            kp, des = orb.detectAndCompute(gray, None)
        if len(_ORB_CACHE) >= _ORB_CACHE_MAX:
            _ORB_CACHE.clear()
        hit = _ORB_CACHE[key] = (kp, des, ("img",) + key)
    return hit


class _KeypointRows:
    """Keypoints of the GPU detector as float32 [n, 6] rows; turns into cv2.KeyPoint objects on demand (drawing)."""

    def __init__(self, rows):
        self.rows = rows
        self._kps = None

    def __len__(self):
        return len(self.rows)

    def as_cv2(self):
        if self._kps is None:
            self._kps = tuple(_sfm.orb.to_cv2_keypoints(self.rows))
        return self._kps


def _cv2_keypoints(kp):
    return kp.as_cv2() if isinstance(kp, _KeypointRows) else kp


class _SlotBank:
    """A small persistent Hamming bank with one slot per recently seen descriptor set, so that the reference's
    N(N-1) pair loop (code/pipeline.py:38-41) uploads every image once and each call is one GPU launch."""

    def __init__(self, n_slots=64, max_feats=1024):
        self.n_slots, self.max_feats = n_slots, max_feats
        self.bank = _sfm.DescriptorBank(n_slots, max_feats, metric="hamming")
        self.slot_of, self.order = {}, []                       # key -> slot, LRU order of keys

    def slot(self, key, des):
        s = self.slot_of.get(key)
        if s is not None:
            self.order.remove(key)
            self.order.append(key)
            return s
        if len(self.order) < self.n_slots:
            s = len(self.order)
        else:
            s = self.slot_of.pop(self.order.pop(0))              # evict the least recently used image
        if isinstance(des, np.ndarray):
            pad = np.zeros((1, self.max_feats, 32), np.uint8)
            pad[0, : len(des)] = des
            self.bank.put(s, pad, counts=[len(des)])
        else:                                                    # descriptors computed on the device (sfm_b200.orb): packed in place
            self.bank.put(s, des.unsqueeze(0))
        self.slot_of[key] = s
        self.order.append(key)
        return s


_SLOTS = None


def _des_key(des):
    return (des.shape, hashlib.blake2b(np.ascontiguousarray(des), digest_size=16).digest())


def match_descriptors_hamming(des1, des2, max_distance=MAX_HAMMING_DISTANCE, _keys=None):
    """The reference's matcher on precomputed binary descriptors -> list[cv2.DMatch].
    Either side empty/None returns [] (the reference returns [] for an empty first image and raises
    cv2.error for an empty second one; pipeline.py only tests truthiness, so [] is superset-safe)."""
    global _SLOTS
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) == 0:
        return []
    on_device = [not isinstance(d, np.ndarray) and hasattr(d, "is_cuda") for d in (des1, des2)]
    if not on_device[0]:
        des1 = np.ascontiguousarray(des1, np.uint8).reshape(len(des1), 32)
    if not on_device[1]:
        des2 = np.ascontiguousarray(des2, np.uint8).reshape(len(des2), 32)
    if any(on_device) and _keys is None:
        raise ValueError("device descriptors need cache keys")
    need = max(len(des1), len(des2))
    if _SLOTS is None or need > _SLOTS.max_feats:
        _SLOTS = _SlotBank(64, max(1024, 2 * need))
    k1, k2 = _keys if _keys is not None else (_des_key(des1), _des_key(des2))
    s1 = _SLOTS.slot(k1, des1)
    s2 = _SLOTS.slot(k2, des2)
    if s1 == s2 and k1 != k2:                                    # (cannot happen with >= 2 slots; keep the invariant explicit)
        raise RuntimeError("descriptor slot collision")
    q, t, d = _sfm.match_pairs_hamming(_SLOTS.bank, [[s1, s2]], max_distance).to_host()[0]
    return [cv2.DMatch(int(a), int(b), 0, float(c)) for a, b, c in zip(q, t, d)]


def match_descriptors_l2(des1, des2, ratio=0.75, ratio_mode="cv2_f32", mutual=False):
    """North-star SIFT workload on two descriptor sets: BFMatcher(NORM_L2).knnMatch(k=2) + Lowe ratio
    -> list[cv2.DMatch] in ascending queryIdx with cv2's float32 distances."""
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) == 0:
        return []
    bank = _sfm.build_bank([des1, des2], metric="l2")
    q, t, d = _sfm.match_pairs(bank, [[0, 1]], ratio=ratio, ratio_mode=ratio_mode, mutual=mutual, with_corr=False).to_host()[0]
    bank.destroy()
    dist = np.sqrt(d.astype(np.float32))
    return [cv2.DMatch(int(a), int(b), 0, float(c)) for a, b, c in zip(q, t, dist)]


def extract_and_match_draw(gray1, gray2):
    """code/feature_matching.py:15-37: extract_and_match plus cv2.drawMatches and a blocking plt.show()."""
    _check_image(gray1, "gray1")
    _check_image(gray2, "gray2")
    kp1, des1, key1 = _extract(gray1)
    kp2, des2, key2 = _extract(gray2)
    cropped_matches = match_descriptors_hamming(des1, des2, _keys=(key1, key2))
    imgDebug = cv2.drawMatches(gray1, _cv2_keypoints(kp1), gray2, _cv2_keypoints(kp2), cropped_matches, None,
                               flags=cv2.DrawMatchesFlags_NOT_DRAW_SINGLE_POINTS)
    plt.imshow(imgDebug), plt.show()
    return cropped_matches


def extract_and_match(gray1, gray2):
    """code/feature_matching.py:41-60: ORB both images, Hamming cross-check match, sort by distance,
    keep the prefix with distance < 26.  Returns a fresh list[cv2.DMatch] (falsy when empty)."""
    _check_image(gray1, "gray1")
    _check_image(gray2, "gray2")
    kp1, des1, key1 = _extract(gray1)
    kp2, des2, key2 = _extract(gray2)
    return match_descriptors_hamming(des1, des2, _keys=(key1, key2))
