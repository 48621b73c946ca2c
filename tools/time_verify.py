"""Times the verification-stage kernels alone on one B200 (CUDA events on the launching stream, L2 not relevant: the
correspondences of a pair are staged in shared memory once):  RANSAC-F (7/8-point), RANSAC-H, two-view pose.
Shapes: the bench's (1,225 pairs x ~2,730 correspondences, mostly inliers, adaptive stop) and config 5's
(4,096 correspondences at 50 % outliers, a fixed hypothesis budget).   python tools/time_verify.py [--pairs5 N] [--hyp N]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import sfm_b200  # noqa: E402
from sfm_b200 import ransac as rs  # noqa: E402
from sfm_b200 import synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def batch(n_pairs, M, outl, seed0, gen=synth.two_view_correspondences, n_unique=50):
    cs = np.zeros((n_unique, M, 4), np.float32)
    for k in range(n_unique):
        p1, p2, _, _ = gen(M, outlier_frac=outl, seed=seed0 + k)
        cs[k, :, :2], cs[k, :, 2:] = p1, p2
    idx = torch.arange(n_pairs) % n_unique
    corr = torch.from_numpy(cs).cuda()[idx.cuda()].contiguous()
    return corr, torch.full((n_pairs,), M, dtype=torch.int32, device="cuda")


def main():
    arg = lambda k, d: int(sys.argv[sys.argv.index(k) + 1]) if k in sys.argv else d  # noqa: E731
    n5, H5 = arg("--pairs5", 1184), arg("--hyp", 10000)
    torch.cuda.set_device(0)
    out = {}
    # bench shape: 1,225 pairs, 2,730 correspondences, ~0.5 % outliers -> one batch of 128 hypotheses
    corr, counts = batch(1225, 2730, 0.005, 100)
    pid = np.arange(1225)
    for solver in ("8pt", "7pt"):
        ms, vb = timed(lambda: rs.verify_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=2000, solver=solver, seed=1, pair_id=pid))
        out[f"ransac_f_{solver}_bench_shape"] = {"ms": ms, "mean_hyp": float(vb.iters.float().mean()), "pairs_per_s": 1225 / ms * 1e3}
    for solver in ("8pt", "7pt"):
        ms, vl = timed(lambda: rs.verify_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=2000, solver=solver, seed=1, pair_id=pid, lo=True))
        out[f"ransac_f_{solver}_bench_shape_lo"] = {"ms": ms, "mean_hyp": float(vl.iters.float().mean()), "mean_inliers": float(vl.n_inliers.float().mean()),
                                                    "mean_inliers_without_lo": float(vb.n_inliers.float().mean())}
    ms, hb = timed(lambda: rs.verify_h_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=2000, seed=1, pair_id=pid))
    out["ransac_h_bench_shape_nonplanar"] = {"ms": ms, "mean_hyp": float(hb.iters.float().mean())}
    tgt = (vb.n_inliers.to(torch.float32) * 0.8).to(torch.int32)
    ms, hb = timed(lambda: rs.verify_h_corr(corr, counts, thr=3.0, confidence=0.99, max_iters=2000, seed=1, pair_id=pid, stop_target=tgt))
    out["ransac_h_bench_shape_nonplanar_stop_target_0.8nF"] = {"ms": ms, "mean_hyp": float(hb.iters.float().mean())}
    cam = rs.camera_rows(synth.K_INTR, None, 1225)
    ms, pb = timed(lambda: rs.recover_pose_corr(corr, counts, vb.F, cam, mask=vb.mask))
    tri = float(vb.n_inliers.sum()) * 5
    out["two_view_pose_bench_shape"] = {"ms": ms, "pairs_per_s": 1225 / ms * 1e3, "triangulations_per_s": tri / ms * 1e3,
                                        "mean_in_front": float(pb.n_good.float().mean())}
    pc, pcounts = batch(1225, 2730, 0.3, 300, gen=synth.planar_correspondences)
    ms, hb = timed(lambda: rs.verify_h_corr(pc, pcounts, thr=3.0, confidence=0.995, max_iters=2000, lo=True, seed=1, pair_id=pid))
    out["ransac_h_planar_30pct_outliers_lo"] = {"ms": ms, "mean_hyp": float(hb.iters.float().mean()), "mean_inliers": float(hb.n_inliers.float().mean())}
    # config-5 shape on a sample of pairs (4 full waves of 296 CTA slots by default)
    corr5, counts5 = batch(n5, 4096, 0.5, 5000)
    pid5 = np.arange(n5)
    for solver in ("8pt", "7pt"):
        ms, vb = timed(lambda: rs.verify_corr(corr5, counts5, thr=3.0, confidence=1.0, max_iters=H5, solver=solver, lo=True, seed=7, pair_id=pid5), reps=2)
        gb = n5 * float(H5) * 4096 * 16 / 1e9
        out[f"ransac_f_{solver}_config5_shape"] = {"ms": ms, "pairs": n5, "hyp": H5, "hypotheses_per_s": n5 * H5 / ms * 1e3,
                                                   "algorithmic_GBps_H_M_16": gb / ms * 1e3, "extrapolated_ms_19900_pairs": ms * 19900 / n5}
    ms, hb = timed(lambda: rs.verify_h_corr(corr5, counts5, thr=3.0, confidence=1.0, max_iters=H5, lo=True, seed=7, pair_id=pid5), reps=2)
    out["ransac_h_config5_shape"] = {"ms": ms, "hypotheses_per_s": n5 * H5 / ms * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
