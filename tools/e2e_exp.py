"""End-to-end variants of the bench step (host buffers in and out), timed with CUDA events + host clock:
bank.put + match_and_verify(fetch='view') against match_and_verify_host with the upload chunked on a side stream."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import sfm_b200  # noqa: E402
from sfm_b200 import synth  # noqa: E402

sc = synth.make_scene(50, 8192, seed=2001)
pairs = synth.exhaustive_pairs(50)
bank = sfm_b200.DescriptorBank(50, 8192)
desc_pin, xy_pin = torch.from_numpy(sc.desc).pin_memory(), torch.from_numpy(sc.xy).pin_memory()
R = dict(ratio=0.75, thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", score="sym_epipolar", lo=False, seed=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=8):
    for _ in range(3):
        fn(); flush.zero_()
    torch.cuda.synchronize()
    ev, host = [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); e1.synchronize()
        host.append(1e3 * (time.perf_counter() - t0)); ev.append(e0.elapsed_time(e1))
        flush.zero_(); torch.cuda.synchronize()
    return round(float(np.median(ev)), 3), round(float(np.median(host)), 3)


def plain(pb=2048):
    bank.put(0, desc_pin, xy=xy_pin)
    return sfm_b200.match_and_verify(bank, pairs, pair_batch=pb, fetch="view", **R)


def put_only():
    bank.put(0, desc_pin, xy=xy_pin)


def resident():
    return sfm_b200.match_and_verify(bank, pairs, **R)


def resident_fetch():
    return sfm_b200.match_and_verify(bank, pairs, fetch="view", **R)


print("put only (H2D + pack)", timed(put_only))
print("resident, no fetch", timed(resident))
print("resident, fetch=view", timed(resident_fetch))
print("put + match_and_verify(fetch=view)", timed(plain))
for nc in (2, 3, 5):
    print(f"match_and_verify_host n_chunks={nc}", timed(lambda: sfm_b200.match_and_verify_host(desc_pin, xy_pin, pairs, bank=bank, n_chunks=nc, fetch="view", **R)))
for nc, pb, g in ((3, 512, 1.0), (3, 512, 1.5), (3, 512, 2.0), (4, 512, 1.5), (4, 512, 2.0), (4, 640, 2.0), (3, 640, 1.5)):
    print(f"match_and_verify_host n_chunks={nc} pair_batch={pb} chunk_growth={g}",
          timed(lambda: sfm_b200.match_and_verify_host(desc_pin, xy_pin, pairs, bank=bank, n_chunks=nc, fetch="view", pair_batch=pb, chunk_growth=g, **R)))
for nc, pb in ((2, 512), (1, 2048)):
    print(f"match_and_verify_host n_chunks={nc} pair_batch={pb}",
          timed(lambda: sfm_b200.match_and_verify_host(desc_pin, xy_pin, pairs, bank=bank, n_chunks=nc, fetch="view", pair_batch=pb, **R)))
