"""Where the time of one ORB extraction goes (1080p): host timeline with a synchronize after every stage, then the plain call.
    python tools/prof_orb.py            (under ncu for the launch list: ncu --metrics gpu__time_duration.sum ... python tools/prof_orb.py once)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from sfm_b200 import orb  # noqa: E402


def textured(rng, h, w):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2), dtype=np.uint8)
    import cv2
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.int32)
    img += rng.integers(-20, 21, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    once = len(sys.argv) > 1 and sys.argv[1] == "once"
    rng = np.random.default_rng(3)
    img = textured(rng, 1080, 1920)
    e = orb.OrbExtractor(1920, 1080)
    for _ in range(1 if once else 3):
        kp, d = e.detect_and_compute(img)
    torch.cuda.synchronize()
    if once:
        return
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kp, d = e.detect_and_compute(img)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"detect_and_compute 1080p: median {np.median(ts):.3f} ms, min {min(ts):.3f} ms, {len(kp)} keypoints")
    # stage split (a synchronize after every stage)
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e.pyramid(img); torch.cuda.synchronize(); t1 = time.perf_counter()
        kp = e.detect(img, _pyramid_done=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"pyramid (upload + 7 resizes + 8 blurs) {1e3 * (t1 - t0):.3f} ms, detect {1e3 * (t2 - t1):.3f} ms")
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        e.detect_and_compute(img)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
