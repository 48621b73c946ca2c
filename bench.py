#!/usr/bin/env python
"""bench.py -- verified pairs/s of the matching + verification hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (tcgen05 match -> ratio filter -> RANSAC-F) over one batch of synthetic
input: BASELINE.json configs[1], the 50-image exhaustive run (1,225 pairs x 8192 SIFT-like features per image).
With N > 1 (torchrun, one rank per GPU) every rank holds the bank and processes its own 1,225-pair block of an
N-times longer pair list (weak scaling, no data-path collective; per-pair summaries are gathered on rank 0).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_IMAGES, N_FEATS = 50, 8192
RANSAC = dict(thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", score="sym_epipolar", lo=False, seed=1)
RATIO = 0.75
OPS_PER_PAIR = 2.0 * N_FEATS * N_FEATS * 128            # algorithmic int8 ops
NOMINAL_INT8_TOPS = 4500.0


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return pk, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------ reference arm (cv2 on the host)
def cv2_pairs_per_s(scene, pairs, n_sample):
    """The reference's arithmetic for this workload (BASELINE.json configs[0]): cv2 BFMatcher L2 knnMatch(k=2) + ratio +
    findFundamentalMat(FM_RANSAC) with all host threads cv2 uses, on a bounded sample of the pair list."""
    from oracle import cv2_ref

    t0 = time.perf_counter()
    verified = 0
    for i, j in pairs[:n_sample]:
        q, t, F, mask = cv2_ref.verified_pair(scene.desc[i], scene.xy[i], scene.desc[j], scene.xy[j], ratio=RATIO,
                                              thr=RANSAC["thr"], confidence=RANSAC["confidence"], max_iters=RANSAC["max_iters"])
        verified += 1
    dt = time.perf_counter() - t0
    return verified / dt, dt


def run_reference(args):
    import cv2

    from sfm_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = synth.make_scene(8, N_FEATS, seed=2001)                   # same generator / feature count as the GPU arm
    pairs = synth.exhaustive_pairs(8)
    n_sample = 6
    for _ in range(args.warmup):
        cv2_pairs_per_s(scene, pairs, 1)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cv2_pairs_per_s(scene, pairs, n_sample)
        rates.append(r)
        times.append(dt)
    value = float(n_sample * len(times) / sum(times))
    cores = int(cv2.getNumThreads())
    sample = f"{n_sample} pairs of 8192x8192x128 per step (knnMatch k=2 f32 + ratio {RATIO} + findFundamentalMat FM_RANSAC 3.0/0.99/2000)"
    line = {
        "impl": "reference", "metric": "verified pairs/s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 50-image exhaustive (1,225 pairs) x 8192 feats/img + RANSAC F; bounded sample", "sample": sample,
                   "cv2": cv2.__version__, "host_cpus": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import sfm_b200
    from sfm_b200 import dist as sdist
    from sfm_b200 import matcher, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: one synthetic scene; rank 0 packs the bank, the others receive it over NVLink (once, untimed)
    scene = synth.make_scene(N_IMAGES, N_FEATS, seed=2001)
    pairs_one = synth.exhaustive_pairs(N_IMAGES)                       # 1,225 pairs
    pairs_all = np.concatenate([pairs_one] * world)                   # weak scaling: 1,225 pairs per rank
    mine = sdist.partition(len(pairs_all), rank, world, "block")
    bank = sfm_b200.DescriptorBank(N_IMAGES, N_FEATS, device=dev)
    if rank == 0:
        bank.put(0, scene.desc, xy=scene.xy)
    sdist.broadcast_bank(bank, src=0)
    torch.cuda.synchronize()
    desc_pin = torch.from_numpy(scene.desc).pin_memory()
    xy_pin = torch.from_numpy(scene.xy).pin_memory()
    my_pairs = pairs_all[mine]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_resident():
        res = sfm_b200.match_and_verify(bank, my_pairs, ratio=RATIO, pair_ids=mine, **RANSAC)
        if world > 1:
            sdist.gather_pair_results({"n_matches": res.n_matches, "n_inliers": res.n_inliers, "F": res.F}, mine, len(pairs_all), 0)
        return res

    def step_e2e():
        bank.put(0, desc_pin, xy=xy_pin)                               # H2D from pinned host memory + pack kernel
        res = sfm_b200.match_and_verify(bank, my_pairs, ratio=RATIO, pair_ids=mine, **RANSAC)
        host = res.to_host(with_matches=True)                          # D2H of everything a caller consumes
        return host

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
            flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms, out = [], None
        l0 = sfm_b200.launch_count()
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            flush.zero_()                                              # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # max over ranks
        return float(t.item()), out, sfm_b200.launch_count() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, res, launches_per_run = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = len(pairs_all) * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public API with host buffers (H2D + pack + match + verify + D2H every step)
    e2e_ms, host, _ = timed(step_e2e, max(1, min(args.steps, 3)), 1)
    e2e_steps = max(1, min(args.steps, 3))
    e2e_value = len(pairs_all) * e2e_steps / (e2e_ms * 1e-3)
    h2d = desc_pin.numel() + xy_pin.numel() * 4 + my_pairs.nbytes + 4 * len(my_pairs) + 4 * N_IMAGES
    d2h = sum(int(v.nbytes) for v in host.values())

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (match_tc_kernel, tensor bound), measured live with CUDA events on the launch stream
        knn = torch.empty((len(my_pairs), bank.feat_stride, 4), dtype=torch.int32, device=dev)
        for _ in range(2):
            sfm_b200.knn2(bank, my_pairs, impl="tcgen05", out=knn, sweep_only=True)
        sw_ms = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sfm_b200.knn2(bank, my_pairs, impl="tcgen05", out=knn, sweep_only=True)      # memset + match_tc_kernel
            e1.record()
            e1.synchronize()
            sw_ms.append(e0.elapsed_time(e1))
        sweep_ms = float(np.mean(sw_ms))
        del knn
        peaks, peak_src = load_peaks()
        achieved = len(my_pairs) * OPS_PER_PAIR / (sweep_ms * 1e-3) / 1e12
        probe_ms, probe_rate = matcher.probe_int8_peak(local, 4096)
        peak = 2.0 * float(peaks["bf16_tflops"])
        # RANSAC scoring: algorithmic bytes H * M * 16 per pair, timed alone on the step's own correspondences
        mb = sfm_b200.match_pairs(bank, my_pairs, ratio=RATIO)
        torch.cuda.synchronize()
        rs_ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            vb = sfm_b200.verify_corr(mb.corr, mb.counts, pair_id=mine, **RANSAC)
            e1.record()
            e1.synchronize()
            rs_ms.append(e0.elapsed_time(e1))
        ransac_ms = float(np.mean(rs_ms))
        m_counts = mb.counts.cpu().numpy().astype(np.float64)
        iters = vb.iters.cpu().numpy().astype(np.float64)
        ransac_bytes = float((iters * m_counts * 16.0).sum())
        ninl = res.n_inliers.cpu().numpy()

        # ---- CPU baseline: the reference's cv2 path on this box's host cores, bounded sample
        import cv2

        cpu_value, cpu_dt = cv2_pairs_per_s(scene, pairs_one, 16)
        line = {
            "metric": "verified pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic",
            "config": {
                "workload": "configs[1]: 50-image exhaustive matching (1,225 pairs) x 8192 features/image + RANSAC F verification",
                "pairs_per_rank": int(len(my_pairs)), "pairs_total": int(len(pairs_all)), "features_per_image": N_FEATS,
                "ratio": RATIO, "ransac": RANSAC, "l2": "flushed between timed iterations (256 MiB write; bank 66 MiB < 126 MB L2)",
                "parallelism": f"pair-sharded x{world}, bank broadcast once (untimed), per-pair summaries gathered on rank 0",
                "mean_matches_per_pair": float(m_counts.mean()), "mean_inliers_per_pair": float(ninl.mean()),
                "mean_hypotheses_per_pair": float(iters.mean()),
            },
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches_per_run),
            "clocks": clocks,
            "roofline": {
                "kernel": "sfm::match_tc_kernel (tcgen05 kind::i8 sweep)", "bound": "tensor", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": f"2 x bf16_tflops of MEASURED_PEAKS.json ({peak_src}; the file has no int8 entry, int8 dense = 2 x bf16 dense)",
                "launch_ms": sweep_ms, "algorithmic_ops_per_launch": len(my_pairs) * OPS_PER_PAIR,
                "frac_of_nominal_4500": achieved / NOMINAL_INT8_TOPS,
                "mma_only_probe_tops": probe_rate / 1e12, "frac_of_mma_only_probe": achieved / (probe_rate / 1e12),
                "matcher_pairs_per_s_sweep_only": len(my_pairs) / (sweep_ms * 1e-3),
            },
            "roofline_ransac": {
                "kernel": "sfm::ransac_f_kernel", "bound": "hbm", "achieved": ransac_bytes / (ransac_ms * 1e-3) / 1e9,
                "peak": float(peaks["hbm_gbs"]), "unit": "GB/s", "frac": ransac_bytes / (ransac_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]),
                "traffic": None, "launch_ms": ransac_ms,
                "note": "algorithmic bytes = hypotheses x correspondences x 16 B; correspondences are staged in shared memory, so DRAM traffic is ~M*16 B per pair and this ratio can exceed 1 (SURVEY.md §8d caveat)",
            },
            "cpu_baseline": {"value": cpu_value, "unit": "pairs/s", "cores": int(cv2.getNumThreads()), "kind": "reference",
                             "sample": f"16 pairs of the same scene through cv2 {cv2.__version__} (knnMatch k=2 on f32 + ratio + findFundamentalMat FM_RANSAC), {cpu_dt:.1f} s"},
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
