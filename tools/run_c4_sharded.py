"""BASELINE.json configs[3] on N GPUs: 1000-image windowed matching (window 20 -> 19,790 pairs) x 32,768 features per image,
pair-sharded, whole result gathered on rank 0 (launch with torchrun; one rank per GPU).  A parity-test case, not a bench
line: prints one JSON line with the timing and the properties checked.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/run_c4_sharded.py [--quick]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sfm_b200  # noqa: E402
from sfm_b200 import dist as sdist  # noqa: E402
from sfm_b200 import synth  # noqa: E402

QUICK = "--quick" in sys.argv


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_img, n_feats, window = (120, 32768, 20) if QUICK else (1000, 32768, 20)
    pairs = synth.windowed_pairs(n_img, window)
    t0 = time.time()
    if rank == 0:
        from tools import run_configs                                   # the GPU scene generator of the single-GPU run (device 0)

        bank, point = run_configs.gpu_scene(n_img, n_feats, shared=16384, stride=400, seed=4001)
    else:
        bank = sfm_b200.DescriptorBank(n_img, n_feats, device=dev)
    torch.cuda.synchronize()
    setup = time.time() - t0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sdist.broadcast_bank(bank, src=0)
    torch.cuda.synchronize()
    bcast_ms = 1e3 * (time.perf_counter() - t0)
    kw = dict(ratio=0.75, thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", seed=1, pair_batch=512)

    def step():
        return sdist.match_and_verify_sharded(bank, pairs, mode="block", gather="full", **kw)

    out, _ = step()                                                     # warm-up: plans, region, NCCL connections
    torch.cuda.synchronize()
    ms = []
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = step()
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    line = None
    if rank == 0:
        nm, ni = out["n_matches"].cpu().numpy(), out["n_inliers"].cpu().numpy()
        gap = pairs[:, 1] - pairs[:, 0]
        expect = 16384 - 400 * gap                                      # scene points both images observe
        assert (nm > 0.9 * expect).all() and (nm < 1.05 * expect + 200).all(), "match counts do not follow the overlap"
        assert (ni > 0.8 * nm).mean() > 0.99
        # a sample of pairs from every rank's block against rank 0's own recomputation: summaries, rows, flags
        lay = sdist.layout(len(pairs), world, "block")
        rng = np.random.default_rng(3)
        sample = np.unique(np.concatenate([rng.choice(o, size=min(4, len(o)), replace=False) for o in lay.owned if len(o)]))
        ref = sfm_b200.match_and_verify(bank, pairs[sample], pair_ids=sample, fetch=True, **kw).to_host()
        ok = all(np.array_equal(out[k][torch.as_tensor(sample, device=dev)].cpu().numpy(), ref[k]) for k in ("n_matches", "n_inliers", "iters", "F"))
        start = out["row_start"][torch.as_tensor(sample, device=dev)].cpu().numpy()
        rows = 0
        for k in range(len(sample)):
            a, b = int(ref["offsets"][k]), int(ref["offsets"][k + 1])
            ok &= np.array_equal(out["matches"][int(start[k]): int(start[k]) + (b - a)].cpu().numpy(), ref["matches"][a:b])
            ok &= np.array_equal(out["inlier"][int(start[k]): int(start[k]) + (b - a)].cpu().numpy(), ref["inlier"][a:b])
            rows += b - a
        assert ok, "sharded result differs from the single-GPU recomputation"
        ops = 2.0 * n_feats * n_feats * 128 * len(pairs)
        best = min(ms)
        line = {"config": "configs[3]: %d-image windowed (window %d), %d pairs x %d feats, pair-sharded over %d GPUs, full gather on rank 0"
                          % (n_img, window, len(pairs), n_feats, world),
                "n_gpus": world, "ms": best, "pairs_per_s": len(pairs) / best * 1e3, "algorithmic_TOPs_whole_job": ops / best / 1e9,
                "bank_GiB": bank.storage.numel() / 2 ** 30, "bank_broadcast_ms": bcast_ms, "transport": out.get("transport"),
                "mean_matches": float(nm.mean()), "mean_inliers": float(ni.mean()), "total_match_rows": int(nm.sum()),
                "selfcheck": {"pairs": int(len(sample)), "rows": int(rows), "equal": True}, "scene_setup_s": setup}
    if world > 1:
        dist.barrier()
        for reg in bank.__dict__.get("_gather_regions", {}).values():
            reg.close()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
