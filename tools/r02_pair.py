"""Round-2 bring-up of the CTA-pair sweep (impl="cluster"): bit-exactness against the one-CTA kernel, then timings.
usage: python tools/r02_pair.py [n_pairs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import sfm_b200
from sfm_b200 import synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1225
dev = torch.device("cuda", 0)
sc = synth.make_scene(50, 8192, seed=2001)
bank = sfm_b200.DescriptorBank(50, 8192)
bank.put(0, sc.desc, xy=sc.xy)
pairs = synth.exhaustive_pairs(50)[:n_pairs]
torch.cuda.synchronize()

# ---- correctness: candidate records may differ in form, the refined kNN table may not
small = pairs[:37]
ref = sfm_b200.knn2(bank, small, impl="tcgen05")
torch.cuda.synchronize()
t0 = time.time()
got = sfm_b200.knn2(bank, small, impl="cluster")
torch.cuda.synchronize()
same = bool(torch.equal(ref, got))
print(f"pair kernel == one-CTA kernel on {len(small)} pairs: {same}  ({time.time() - t0:.2f} s)", flush=True)
if not same and os.environ.get("SFM_PAIR_TIMING_ONLY"):
    print("  (timing only: results are expected to differ)")
elif not same:
    d = (ref != got).any(dim=2)
    bad = d.nonzero()
    print("  differing rows:", int(d.sum()), "first:", bad[:5].tolist())
    for p_, r_ in bad[:3].tolist():
        print("   ref", ref[p_, r_].tolist(), "got", got[p_, r_].tolist())
    sys.exit(1)
ragged = sfm_b200.build_bank([sc.desc[0][:5000], sc.desc[1][:3001], sc.desc[2][:77], sc.desc[3]])
rp = [[0, 1], [1, 0], [2, 3], [3, 2], [1, 2], [0, 3]]
a, b = sfm_b200.knn2(ragged, rp, impl="tcgen05"), sfm_b200.knn2(ragged, rp, impl="cluster")
print("ragged images equal:", bool(torch.equal(a, b)), flush=True)


def timeit(impl, sweep_only, reps=5):
    knn = torch.empty((len(pairs), bank.feat_stride, 4), dtype=torch.int32, device=dev)
    for _ in range(2):
        sfm_b200.knn2(bank, pairs, impl=impl, out=knn, sweep_only=sweep_only)
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sfm_b200.knn2(bank, pairs, impl=impl, out=knn, sweep_only=sweep_only)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


ops = 2.0 * 8192 * 8192 * 128 * len(pairs)
for name, impl, so in (("one-CTA sweep", "tcgen05", 4), ("pair sweep", "cluster", 4), ("pair, K-extension MMA only (epilogue bound)", "cluster", 5),
                       ("pair, epilogue releases at once (MMA bound)", "cluster", 6)):
    ms = timeit(impl, so)
    cyc = ms * 1e-3 * 1.965e9 / (len(pairs) * 32 * 64 / 148)
    print(f"{name:48s} {ms:8.3f} ms  {ops / ms / 1e9:8.1f} TOP/s  {ops / ms / 1e9 / 4500:.3f} of nominal  ~{cyc:.0f} cycles per B tile @1965", flush=True)
