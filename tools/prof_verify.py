"""Short deterministic verification-stage run for ncu: RANSAC-F (8-point then 7-point) on 296 pairs x 4096
correspondences at 50 % outliers, 1024 hypotheses each (config-5 shape), then RANSAC-H and the pose kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import numpy as np
import torch

from sfm_b200 import ransac as rs
from sfm_b200 import synth
from tools.time_verify import batch

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 2
corr, counts = batch(296, 4096, 0.5, 5000, n_unique=16)
pid = np.arange(296)
cam = rs.camera_rows(synth.K_INTR, None, 296)
for _ in range(n_rep):
    v8 = rs.verify_corr(corr, counts, thr=3.0, confidence=1.0, max_iters=1024, solver="8pt", lo=False, seed=7, pair_id=pid)
    v7 = rs.verify_corr(corr, counts, thr=3.0, confidence=1.0, max_iters=1024, solver="7pt", lo=False, seed=7, pair_id=pid)
    vh = rs.verify_h_corr(corr, counts, thr=3.0, confidence=1.0, max_iters=1024, lo=True, seed=7, pair_id=pid)
    pb = rs.recover_pose_corr(corr, counts, v8.F, cam, mask=v8.mask)
torch.cuda.synchronize()
print("ok", int(v8.n_inliers.sum()), int(v7.n_inliers.sum()), int(vh.n_inliers.sum()), int(pb.n_good.sum()))
