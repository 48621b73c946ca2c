"""Same-box A/B of the two-stream pipeline (overlap=True / False) through the public API: resident one-batch job, resident
multi-batch job, and the host-buffer path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import numpy as np
import torch

import sfm_b200
from sfm_b200 import synth

R = dict(thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", score="sym_epipolar", lo=False, seed=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


for n_img, batch in ((50, 2048), (100, 2048), (100, 1024)):
    sc = synth.make_scene(n_img, 8192, seed=2001)
    pairs = synth.exhaustive_pairs(n_img)
    bank = sfm_b200.DescriptorBank(n_img, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    desc, xy = torch.from_numpy(sc.desc).pin_memory(), torch.from_numpy(sc.xy).pin_memory()
    for rep in range(2):
        row = []
        for ov in (False, True):
            row.append(timeit(lambda: sfm_b200.match_and_verify(bank, pairs, ratio=0.75, pair_batch=batch, overlap=ov, **R)))
        for ov in (False, True):
            row.append(timeit(lambda: sfm_b200.match_and_verify_host(desc, xy, pairs, bank=bank, n_chunks=3, ratio=0.75, pair_batch=512, fetch="view", overlap=ov, **R), reps=7))
        print(f"{len(pairs)} pairs, batch {batch}: resident serial {row[0]:.3f} overlapped {row[1]:.3f} ms | host path serial {row[2]:.3f} overlapped {row[3]:.3f} ms", flush=True)
    del bank
