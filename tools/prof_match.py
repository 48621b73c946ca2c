"""Short, deterministic matcher run for ncu: 8 images x 8192 features, 56 pairs, tcgen05 sweep + refine."""
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
sc = synth.make_scene(8, 8192, seed=1)
bank = sfm_b200.DescriptorBank(8, 8192)
bank.put(0, sc.desc, xy=sc.xy)
pairs = np.concatenate([synth.exhaustive_pairs(8)] * 2)      # 56 pairs ~ 12 units per SM
out = torch.empty((len(pairs), bank.feat_stride, 4), dtype=torch.int32, device="cuda")
for _ in range(n_rep):
    sfm_b200.knn2(bank, pairs, impl="tcgen05", out=out)
torch.cuda.synchronize()
print("ok", int(out[0, 0, 0]))
