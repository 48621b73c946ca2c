"""Candidates the refinement recomputes per pair, fused form (early stop after the first tile) against the kNN-table form."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import sfm_b200  # noqa: E402
from sfm_b200 import matcher, synth  # noqa: E402

sc = synth.make_scene(50, 8192, seed=2001)
pairs = synth.exhaustive_pairs(50)[:400]
bank = sfm_b200.DescriptorBank(50, 8192)
bank.put(0, sc.desc, xy=sc.xy)
prev = matcher.refine_stats(True)
for fused in (True, False):
    counts, offsets, matches, corr = matcher.match_pairs_packed(bank, pairs, ratio=0.75, fused=fused)
    torch.cuda.synchronize()
    cur = matcher.refine_stats(True)
    print("fused" if fused else "table", "brute-forced rows", cur[0] - prev[0], "candidates per pair", (cur[1] - prev[1]) / len(pairs),
          "matches per pair", int(counts.sum()) / len(pairs))
    prev = cur
matcher.refine_stats(False)
