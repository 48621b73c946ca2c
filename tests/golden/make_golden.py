"""Generate the committed golden fixtures.  Run in the BUILD container only
(needs /root/reference and cv2):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so the
fixtures are outputs of the reference's own function and of the cv2 calls it
is built on, captured here:

* ref_orb_hamming.npz  -- three synthetic grayscale images, their cv2 ORB
  keypoints/descriptors, and the output of the REFERENCE function
  ``extract_and_match`` (code/feature_matching.py:41-60, imported unmodified from
  /root/reference/code) for all six ordered pairs.
* cv2_l2_knn.npz       -- cv2.BFMatcher(NORM_L2).knnMatch(k=2) + the Python ratio
  idiom + crossCheck on SIFT-like descriptors with planted matches, duplicate rows
  (tie-breaks) and ratio-boundary cases.
* cv2_fm_ransac.npz    -- cv2.findFundamentalMat(FM_RANSAC) F and mask on a
  synthetic two-view scene.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))

import cv2  # noqa: E402

from oracle import cv2_ref  # noqa: E402
from sfm_b200 import synth  # noqa: E402


def textured_images(seed=7, n=3, w=320, h=240):
    rng = np.random.default_rng(seed)
    big = np.zeros((h * 2, w * 2), np.float32)
    for _ in range(900):
        cx, cy = rng.integers(0, w * 2), rng.integers(0, h * 2)
        r = int(rng.integers(3, 14))
        col = float(rng.uniform(40, 255))
        if rng.random() < 0.5:
            cv2.circle(big, (int(cx), int(cy)), r, col, -1)
        else:
            cv2.rectangle(big, (int(cx - r), int(cy - r)), (int(cx + r), int(cy + r)), col, -1)
    big = cv2.GaussianBlur(big, (0, 0), 1.0)
    imgs = []
    for k in range(n):
        ang = 4.0 * k
        M = cv2.getRotationMatrix2D((w, h), ang, 1.0 + 0.03 * k)
        M[:, 2] += (-w / 2 + 6 * k, -h / 2 - 4 * k)
        im = cv2.warpAffine(big, M, (w, h))
        im = im + rng.normal(0, 2.0, im.shape)
        imgs.append(np.clip(im, 0, 255).astype(np.uint8))
    return np.stack(imgs)


def make_ref_orb_hamming():
    ref = cv2_ref.import_reference_feature_matching()
    imgs = textured_images()
    out = {"images": imgs}
    for k, im in enumerate(imgs):
        kp, des = cv2_ref.orb_extract(im)
        out[f"des{k}"] = des
        out[f"xy{k}"] = np.array([p.pt for p in kp], np.float32)
    for i in range(len(imgs)):
        for j in range(len(imgs)):
            if i == j:
                continue
            m = ref.extract_and_match(imgs[i], imgs[j])      # the reference function itself
            q, t, d = cv2_ref.dmatches_to_arrays(m)
            out[f"q_{i}_{j}"], out[f"t_{i}_{j}"], out[f"d_{i}_{j}"] = q, t, d
            print(f"reference extract_and_match({i},{j}): {len(q)} matches")
    np.savez_compressed(os.path.join(HERE, "ref_orb_hamming.npz"), **out)


def make_cv2_l2_knn():
    rng = np.random.default_rng(11)
    B = synth.sift_like(rng, 400)
    A = synth.sift_like(rng, 300)
    A[:150] = synth.observe(rng, B[rng.permutation(400)[:150]], 6.0)   # planted matches
    B[300] = B[5]; B[377] = B[5]                                        # triplicated train row
    A[200] = B[5]                                                       # exact hit -> D1=D2=0 tie
    A[201] = B[17]; B[18] = B[17]                                       # duplicate neighbours
    # ratio boundary: craft a query whose two best distances are D1=18, D2=32 (16*18 == 9*32)
    B[390] = 0; B[391] = 0; A[202] = 0
    B[390, :18] = 1          # D1 = 18
    B[391, :32] = 1          # D2 = 32
    A[203] = 0; A[203, 100:] = 200
    B[392] = A[203]; B[393] = A[203]
    B[392, :27] = 1          # D1 = 27
    B[393, :48] = 1          # D2 = 48
    idx1, d1, idx2, d2 = cv2_ref.l2_knn2(A, B)
    q, t, d = cv2_ref.l2_ratio_match(A, B, 0.75)
    cq, ct, cd = cv2_ref.l2_crosscheck(A, B)
    np.savez_compressed(
        os.path.join(HERE, "cv2_l2_knn.npz"), A=A, B=B, idx1=idx1, d1=d1, idx2=idx2, d2=d2,
        ratio_q=q, ratio_t=t, ratio_d=d, cross_q=cq, cross_t=ct, cross_d=cd, cv2_version=cv2.__version__,
    )
    print(f"cv2 knn: {len(idx1)} rows, ratio keeps {len(q)}, crossCheck keeps {len(cq)}")


def make_cv2_fm_ransac():
    p1, p2, gt, Ft = synth.two_view_correspondences(400, outlier_frac=0.3, seed=21)
    F, mask = cv2_ref.find_fundamental(p1, p2, 3.0, 0.99, 2000)
    np.savez_compressed(os.path.join(HERE, "cv2_fm_ransac.npz"), pts1=p1, pts2=p2, gt=gt, F_true=Ft, F=F, mask=mask,
                        cv2_version=cv2.__version__)
    print(f"cv2 FM_RANSAC: {int(mask.sum())} inliers of {len(mask)} (gt {int(gt.sum())})")


if __name__ == "__main__":
    make_ref_orb_hamming()
    make_cv2_l2_knn()
    make_cv2_fm_ransac()
