"""The reference's per-pair call through the drop-in on cached 1080p images: time per call and a cProfile of the host side."""
import os, sys, time, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
    sys.path.insert(0, p)
import cv2, torch
import feature_matching as fm
rng = np.random.default_rng(0)
def textured(h, w):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    return np.clip(img.astype(int) + rng.normal(0, 10, (h, w)).astype(int), 0, 255).astype(np.uint8)
imgs = [textured(1080, 1920) for _ in range(4)]
for a in imgs:
    for b in imgs:
        if a is not b: fm.extract_and_match(a, b)
torch.cuda.synchronize()
t0 = time.perf_counter(); n = 0
for _ in range(5):
    for a in imgs:
        for b in imgs:
            if a is not b: fm.extract_and_match(a, b); n += 1
print("cached per-pair call: %.3f ms" % (1e3 * (time.perf_counter() - t0) / n))
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    for a in imgs:
        for b in imgs:
            if a is not b: fm.extract_and_match(a, b)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
