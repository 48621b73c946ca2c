"""Recover ORB's 256-pair rBRIEF sampling pattern from the installed cv2 with probe images.

The reference's extraction step is ``cv2.ORB_create().detectAndCompute`` (code/feature_matching.py:42-45).  OpenCV's sources
are not on disk, so the pattern (256 test pairs a_i, b_i in [-15, 15]^2; bit i = I(a_i) < I(b_i) on the 7x7 sigma-2 blurred
level image, rotated by the keypoint angle) is measured: ``orb.compute`` keeps a provided keypoint's angle, so with angle 0
the sample positions are the pattern itself, and saturating STEP images locate them absolutely:

* a vertical step (dark left of X0, bright from X0 on) blurs into a ramp that is strictly increasing on [X0-3, X0+3] and flat
  outside; for a pair with a.x < b.x the bit is 1 exactly for X0 in [a.x - 2, b.x + 3]; pairs with a.x > b.x answer to the
  mirrored step; the same along y;
* pairs with a.x == b.x are blind to x-steps: a QUADRANT image (bright where x >= X0 and y >= Y0, Y0 placed at the lower
  point) answers 1 exactly while the common column still receives light, X0 <= x + 3; the same for a.y == b.y.

Writes tests/golden/orb_pattern.npz (int8 [256, 4] = a.x, a.y, b.x, b.y) and sfm-project_b200/csrc/orb_pattern.inc, and
checks the result by recomputing every descriptor of the golden images (tests/golden/ref_orb_hamming.npz) with
oracle/orb_oracle.py.  Run in the build container: python tools/recover_orb_pattern.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N, C, R = 161, 80, 24            # probe image size, keypoint position, scan radius


def bits_of(img):
    orb = cv2.ORB_create()
    kp = [cv2.KeyPoint(float(C), float(C), 31.0, 0.0, 1.0, 0, -1)]
    kp, des = orb.compute(img, kp)
    assert len(kp) == 1 and kp[0].angle == 0.0
    return np.unpackbits(des[0], bitorder="little").astype(bool)        # bit i of the descriptor = byte i // 8, bit i % 8


def step_image(axis, pos, flip, quad=None):
    """bright (255) where coordinate(axis) - C >= pos (or < pos when flip); quad = (other-axis position, flip) ANDs a second half-plane."""
    yy, xx = np.mgrid[0:N, 0:N]
    c = (xx if axis == 0 else yy) - C
    m = (c < pos) if flip else (c >= pos)
    if quad is not None:
        o = (yy if axis == 0 else xx) - C
        m &= (o < quad[0]) if quad[1] else (o >= quad[0])
    return np.where(m, 255, 0).astype(np.uint8)


def scan_axis(axis):
    """(lo, hi) per bit and polarity: the range of step positions for which the bit is 1 (None if never)."""
    out = {}
    for flip in (False, True):
        on = np.array([bits_of(step_image(axis, p, flip)) for p in range(-R, R + 1)])       # [positions, 256]
        rng = []
        for i in range(256):
            idx = np.nonzero(on[:, i])[0]
            if len(idx):
                assert np.array_equal(idx, np.arange(idx[0], idx[-1] + 1)), "bit response is not an interval"
                rng.append((idx[0] - R, idx[-1] - R))
            else:
                rng.append(None)
        out[flip] = rng
    return out


def main():
    pat = np.full((256, 4), 127, np.int64)
    for axis in (0, 1):
        sc = scan_axis(axis)
        for i in range(256):
            up, down = sc[False][i], sc[True][i]
            assert up is None or down is None, "a pair cannot answer to both polarities"
            if up is not None:           # a < b along this axis: bit on for X0 in [a - 2, b + 3]
                pat[i, axis], pat[i, 2 + axis] = up[0] + 2, up[1] - 3
            elif down is not None:       # a > b: bright for c < X0, ramp decreasing; bit on for X0 in [b - 2, a + 3]
                pat[i, 2 + axis], pat[i, axis] = down[0] + 2, down[1] - 3
    # pairs that share a coordinate: quadrant probes along that axis, the other axis' step placed at the larger of the two
    for axis in (0, 1):
        other = 1 - axis
        for i in np.nonzero(pat[:, axis] == 127)[0]:
            ao, bo = pat[i, other], pat[i, 2 + other]
            assert ao != 127 and ao != bo, "degenerate pair"
            # second half-plane: bright on the side of the point that must be BRIGHTER for the bit to be 1 (b), edge at that point
            quad = (max(ao, bo), False) if bo > ao else (min(ao, bo) + 1, True)
            on = np.array([bits_of(step_image(axis, p, False, quad))[i] for p in range(-R, R + 1)])
            idx = np.nonzero(on)[0]
            assert len(idx) and idx[0] == 0, "quadrant probe did not answer"
            pat[i, axis] = pat[i, 2 + axis] = idx[-1] - R - 3
    assert np.abs(pat).max() <= 15, pat[np.abs(pat).max(axis=1) > 15]
    pat = pat.astype(np.int8)
    np.savez(os.path.join(ROOT, "tests", "golden", "orb_pattern.npz"), pattern=pat, cv2_version=cv2.__version__)
    with open(os.path.join(ROOT, "sfm-project_b200", "csrc", "orb_pattern.inc"), "w") as f:
        f.write("// ORB rBRIEF test pairs (a.x, a.y, b.x, b.y), measured from cv2 %s by tools/recover_orb_pattern.py -- do not edit\n" % cv2.__version__)
        for i in range(256):
            f.write("{%d, %d, %d, %d},%s" % (*pat[i], "\n" if i % 8 == 7 else " "))
    print("pattern written; first pairs:", pat[:4].tolist())
    # ---- check: every descriptor of the golden images
    from oracle import orb_oracle

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_orb_hamming.npz"))
    bad = 0
    for k, img in enumerate(g["images"]):
        kp, des = cv2.ORB_create().detectAndCompute(img, None)
        mine = orb_oracle.describe(img, kp, pat)
        bad += int((mine != des).any(axis=1).sum())
        assert np.array_equal(des, g[f"des{k}"])
    print("golden images: descriptors differing from cv2:", bad)
    assert bad == 0


if __name__ == "__main__":
    main()
