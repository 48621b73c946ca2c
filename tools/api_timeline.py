"""Timeline of one match_and_verify call on the bench job (CUPTI through torch.profiler): where the time between the
first and the last device activity goes, including gaps.  Diagnostic."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import sfm_b200  # noqa: E402
from sfm_b200 import synth  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

sc = synth.make_scene(50, 8192, seed=2001)
pairs = synth.exhaustive_pairs(50)
bank = sfm_b200.DescriptorBank(50, 8192)
bank.put(0, sc.desc, xy=sc.xy)
R = dict(ratio=0.75, thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", score="sym_epipolar", lo=False, seed=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    sfm_b200.match_and_verify(bank, pairs, **R)
    flush.zero_()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        sfm_b200.match_and_verify(bank, pairs, **R)
        e1.record()
        t1 = time.perf_counter()
        e1.synchronize()
        print(f"events {e0.elapsed_time(e1):.3f} ms, host enqueue {1e3 * (t1 - t0):.3f} ms")
        flush.zero_()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
# last call: the events after the second-to-last flush kernel
idx = [i for i, e in enumerate(ev) if "FillFunctor<unsigned char>" in e.name or "Memset" in e.name and e.time_range.elapsed_us() > 40]
start = idx[-2] + 1 if len(idx) >= 2 else 0
t_first = ev[start].time_range.start
prev_end = t_first
for e in ev[start:]:
    gap = e.time_range.start - prev_end
    print(f"+{(e.time_range.start - t_first) / 1e3:8.3f} ms  dur {e.time_range.elapsed_us() / 1e3:7.3f} ms  gap {gap / 1e3:7.3f} ms  {e.name[:70]}")
    prev_end = max(prev_end, e.time_range.end)
