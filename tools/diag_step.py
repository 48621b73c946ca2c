"""Host-vs-device timing of the resident step (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))
import numpy as np, torch
import sfm_b200
from sfm_b200 import synth
import bench

sc = synth.make_scene(50, 8192, seed=2001)
pairs = synth.exhaustive_pairs(50)
bank = sfm_b200.DescriptorBank(50, 8192)
bank.put(0, sc.desc, xy=sc.xy)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
R = dict(thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", score="sym_epipolar", lo=False, seed=1)

def run(tag, n=8, sampler=False):
    s = bench.ClockSampler(0)
    if sampler: s.start()
    rows = []
    for it in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        res = sfm_b200.match_and_verify(bank, pairs, ratio=0.75, **R)
        e1.record(); t1 = time.perf_counter()
        e1.synchronize(); t2 = time.perf_counter()
        rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t0), e0.elapsed_time(e1)))
        flush.zero_()
    if sampler: print(tag, "clocks", s.stop())
    for r in rows: print(f"{tag}: host enqueue {r[0]:7.2f} ms  host total {r[1]:7.2f} ms  device {r[2]:7.2f} ms")

run("nosampler")
run("sampler", sampler=True)
run("nosampler2")
