"""Timeline of the CTA-pair sweep (library built with -DSFM_TC2_TRACE=1): prints per-tile event times of cluster 0."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import sfm_b200
from sfm_b200 import _lib, synth

sc = synth.make_scene(4, 8192, seed=2001)
bank = sfm_b200.DescriptorBank(4, 8192)
bank.put(0, sc.desc, xy=sc.xy)
pairs = synth.exhaustive_pairs(4)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(2):
    sfm_b200.knn2(bank, np.concatenate([pairs] * 40), impl="cluster", sweep_only=mode)
torch.cuda.synchronize()
buf = np.zeros(8 * 96 * 6, np.int64)
n = _lib.lib().sfm_debug_pair_trace(buf.ctypes.data_as(C.c_void_p), len(buf))
if n <= 0:
    print("library was not built with -DSFM_TC2_TRACE=1")
    sys.exit(0)
T = buf.reshape(8, 96, 6)
t0 = T[0, 8, 0]
p0 = T[5, 8, 0]
print("tile | issuer: top b_full t_empty issued | epi(set of tile): top c0 all rel nextwait nextfull | producer: top empty issued || peer epi | peer producer")
for t in list(range(8, 20)) + list(range(60, 72)):
    s = t & 1
    iss = (T[0, t, :4] - t0).tolist()
    epi = (T[1 + s, t, :6] - t0).tolist()
    pro = (T[3, t, :3] - t0).tolist()
    pepi = (T[5 + s, t, :6] - p0).tolist()
    ppro = (T[7, t, :3] - p0).tolist()
    print(t, "|", *iss, "|", *epi, "|", *pro, "||", *pepi, "|", *ppro)
d = np.diff(T[0, 8:90, 3])
print("issuer period per tile: mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
