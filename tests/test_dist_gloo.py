"""world_size-2 gloo tests of the pair-sharding plumbing (CPU tensors; the same code runs under NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_every_pair_once():
    from sfm_b200.dist import partition

    for n in (0, 1, 7, 1225, 19900):
        for ws in (1, 2, 4, 8):
            for mode in ("block", "cyclic"):
                got = np.concatenate([partition(n, r, ws, mode) for r in range(ws)])
                assert sorted(got.tolist()) == list(range(n))
    b = partition(10, 1, 4, "block")
    assert b.tolist() == [3, 4, 5]
    with pytest.raises(ValueError):
        partition(10, 0, 2, "zigzag")


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from sfm_b200.dist import gather_pair_results, gather_summaries, gather_varlen, partition
    from types import SimpleNamespace

    n_total = 11
    mine = partition(n_total, rank, ws, "block")
    # per-pair payloads that encode the global pair index
    local = {
        "n_matches": torch.tensor([100 + int(i) for i in mine], dtype=torch.int32),
        "F": torch.stack([torch.full((3, 3), float(i), dtype=torch.float64) for i in mine]) if len(mine) else torch.zeros((0, 3, 3), dtype=torch.float64),
    }
    out = gather_pair_results(local, mine, n_total, dst=0)
    ragged = gather_varlen(torch.arange(rank + 2, dtype=torch.int64) + 10 * rank, dst=0)
    res = SimpleNamespace(n_matches=local["n_matches"], n_inliers=local["n_matches"] - 50, iters=torch.full((len(mine),), 7, dtype=torch.int32),
                          F=local["F"])
    summ = gather_summaries(res, mine, n_total, dst=0)
    cyc = partition(n_total, rank, ws, "cyclic")
    res_c = SimpleNamespace(n_matches=torch.tensor([200 + int(i) for i in cyc], dtype=torch.int32), n_inliers=torch.zeros(len(cyc), dtype=torch.int32),
                            iters=torch.zeros(len(cyc), dtype=torch.int32), F=torch.zeros((len(cyc), 3, 3), dtype=torch.float64))
    summ_c = gather_summaries(res_c, cyc, n_total, dst=0)
    # optional stages present: H / n_inliers_h and R / t / n_pose travel in the same single collective
    res_x = SimpleNamespace(n_matches=local["n_matches"], n_inliers=local["n_matches"] - 50, iters=torch.full((len(mine),), 7, dtype=torch.int32),
                            F=local["F"], H=local["F"] + 0.5, n_inliers_h=local["n_matches"] - 60, R=local["F"] * 2.0,
                            t=torch.stack([torch.tensor([float(i), 1.0, 2.0], dtype=torch.float64) for i in mine]) if len(mine) else torch.zeros((0, 3), dtype=torch.float64),
                            n_pose=local["n_matches"] - 70)
    summ_x = gather_summaries(res_x, mine, n_total, dst=0)
    if rank == 0:
        ok = out["n_matches"].tolist() == [100 + i for i in range(n_total)]
        ok &= summ["n_matches"].tolist() == [100 + i for i in range(n_total)] and summ["n_inliers"].tolist() == [50 + i for i in range(n_total)]
        ok &= summ["iters"].tolist() == [7] * n_total and all(float(summ["F"][i, 2, 1]) == float(i) for i in range(n_total))
        ok &= summ_c["n_matches"].tolist() == [200 + i for i in range(n_total)]
        ok &= all(float(out["F"][i, 0, 0]) == float(i) for i in range(n_total))
        ok &= ragged.tolist() == [0, 1, 10, 11, 12]
        ok &= summ_x["n_inliers_h"].tolist() == [40 + i for i in range(n_total)] and summ_x["n_pose"].tolist() == [30 + i for i in range(n_total)]
        ok &= all(float(summ_x["H"][i, 1, 1]) == i + 0.5 and float(summ_x["R"][i, 0, 2]) == 2.0 * i and summ_x["t"][i].tolist() == [float(i), 1.0, 2.0]
                  for i in range(n_total))
        ok &= "H" not in summ and summ_x["F"].shape == (n_total, 3, 3) and summ_x["t"].shape == (n_total, 3)
        q.put(bool(ok))
    else:
        assert out is None and ragged is None and summ is None and summ_c is None and summ_x is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
