"""Short, deterministic full step for ncu: 12 images x 8192 features, 2 x 66 pairs, match -> filter -> RANSAC-F."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))

import numpy as np
import torch

import sfm_b200
from sfm_b200 import synth

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = synth.make_scene(12, 8192, seed=1)
bank = sfm_b200.DescriptorBank(12, 8192)
bank.put(0, sc.desc, xy=sc.xy)
pairs = np.concatenate([synth.exhaustive_pairs(12)] * 3)      # 198 pairs
for _ in range(n_rep):
    res = sfm_b200.match_and_verify(bank, pairs, ratio=0.75, thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", seed=1)
torch.cuda.synchronize()
print("ok", int(res.n_inliers.sum()))
