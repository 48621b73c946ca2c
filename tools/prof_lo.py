"""ncu target: RANSAC-F with the LO refit on the bench shape (1,225 pairs x 2,730 correspondences, 8-point)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import numpy as np
import torch

from sfm_b200 import ransac as rs
from tools.time_verify import batch

corr, counts = batch(1225, 2730, 0.005, 100)
pid = np.arange(1225)
for _ in range(3):
    v = rs.verify_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", seed=1, pair_id=pid, lo=True)
torch.cuda.synchronize()
print("ok", int(v.n_inliers.sum()))
