"""Multi-GPU parity on hardware: the pair-sharded run on N GPUs returns, on rank 0, exactly what one GPU
computes for the whole pair list -- summaries, packed match rows and inlier flags, for both transports of the row gather
(peer-memory push over NVLink, grouped send/recv) and both partitions.  The reference loop this shards is
code/pipeline.py:36-47.  Skipped when fewer than 2 GPUs are visible."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _worker(rank, ws, port, q):
    import faulthandler

    faulthandler.dump_traceback_later(150, exit=True)                   # a hung collective shows where, and ends the test
    try:
        _worker_body(rank, ws, port, q)
    except Exception:                                                   # the parent must hear about it instead of waiting
        import traceback

        q.put({"error": f"rank {rank}: {traceback.format_exc()}"})
        raise
    finally:
        faulthandler.cancel_dump_traceback_later()


def _worker_body(rank, ws, port, q):
    for p in (ROOT, os.path.join(ROOT, "sfm-project_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
    import sfm_b200
    from sfm_b200 import dist as sdist
    from sfm_b200 import synth

    n_img, n_feat = 7, 2048
    sc = synth.make_scene(n_img, n_feat, seed=77)
    pairs = synth.exhaustive_pairs(n_img)                               # 21 pairs: uneven blocks for 2 ranks
    bank = sfm_b200.DescriptorBank(n_img, n_feat, device=dev)
    if rank == 0:
        bank.put(0, sc.desc, xy=sc.xy)
    sdist.broadcast_bank(bank, src=0)
    prm = dict(ratio=0.75, thr=3.0, confidence=0.99, max_iters=512, solver="8pt", seed=9, lo=True)
    report = {}
    base = None
    if rank == 0:
        base = sfm_b200.match_and_verify(bank, pairs, fetch=True, pair_batch=8, **prm).to_host()
    for transport in ("p2p", "sendrecv"):
        for mode in ("block", "cyclic"):
            for rep in range(3):                                        # three times: regions alternate, the third job reuses the first one (fence) and rewrites it
                out, local = sdist.match_and_verify_sharded(bank, pairs, mode=mode, transport=transport, pair_batch=4, **prm)
            torch.cuda.synchronize()
            if rank == 0:
                ok = out["transport"] == transport
                for k in ("n_matches", "n_inliers", "iters", "F"):
                    ok &= np.array_equal(out[k].cpu().numpy(), base[k])
                start = out["row_start"].cpu().numpy()
                m, inl = out["matches"].cpu().numpy(), out["inlier"].cpu().numpy()
                for p in range(len(pairs)):
                    a, b = base["offsets"][p], base["offsets"][p + 1]
                    ok &= np.array_equal(m[start[p]: start[p] + (b - a)], base["matches"][a:b])
                    ok &= np.array_equal(inl[start[p]: start[p] + (b - a)], base["inlier"][a:b])
                report[f"{transport}/{mode}"] = bool(ok)
    # the optional stages ride along, and a rank with an EMPTY block (1 pair, 2 ranks) contributes a tile of the right width
    K = synth.K_INTR
    out, _ = sdist.match_and_verify_sharded(bank, pairs[:1], homography=True, intrinsics=K, **prm)
    out3, _ = sdist.match_and_verify_sharded(bank, pairs[:5], homography=True, intrinsics=K, mode="cyclic", pair_batch=2, **prm)
    torch.cuda.synchronize()
    if rank == 0:
        b1 = sfm_b200.match_and_verify(bank, pairs[:5], fetch=True, homography=True, intrinsics=K, **prm).to_host()
        ok = True
        for k in ("n_matches", "n_inliers", "F", "H", "n_inliers_h", "R", "t", "n_pose"):
            ok &= np.array_equal(out[k].cpu().numpy(), b1[k][:1]) and np.array_equal(out3[k].cpu().numpy(), b1[k])
        start = out3["row_start"].cpu().numpy()
        for name in ("matches", "inlier", "inlier_h", "in_front", "points3d"):
            got = out3[name].cpu().numpy()
            for p in range(5):
                a, b = b1["offsets"][p], b1["offsets"][p + 1]
                ok &= np.array_equal(got[start[p]: start[p] + (b - a)], b1[name][a:b])
        report["optional stages + empty rank"] = bool(ok)
        empty = sfm_b200.match_and_verify(bank, np.zeros((0, 2), np.int32), homography=True, intrinsics=K)
        report["P == 0 carries H / R"] = empty.H is not None and empty.R is not None and empty.H.shape == (0, 3, 3)
    # sharded upload: every rank packs its slice of the images, the packed sections are all-gathered -> the same bank as one put
    bank2 = sfm_b200.DescriptorBank(8, n_feat, device=dev)               # 7 images over 2 ranks: slices of 4 and 3
    sdist.upload_bank_sharded(bank2, sc.desc, sc.xy)
    torch.cuda.synchronize()
    ref_bank = sfm_b200.DescriptorBank(8, n_feat, device=dev)
    ref_bank.put(0, sc.desc, xy=sc.xy)
    ok = bank2.n_images == n_img
    for name in ("desc", "ext", "norm", "xy"):
        per_image = bank2.section(name).numel() // 8
        ok &= bool(torch.equal(bank2.section(name)[: n_img * per_image], ref_bank.section(name)[: n_img * per_image]))
    ok &= bool(torch.equal(bank2.counts[:n_img], ref_bank.counts[:n_img]))
    h2 = sfm_b200.match_and_verify(bank2, pairs[:6], fetch=True, **prm).to_host()
    h1 = sfm_b200.match_and_verify(ref_bank, pairs[:6], fetch=True, **prm).to_host()
    ok &= all(np.array_equal(h1[k], h2[k]) for k in ("matches", "inlier", "F", "n_inliers"))
    flags = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        report["sharded upload == one put (every rank)"] = bool(flags.item())
        q.put(report)
    dist.barrier()
    for reg in bank.__dict__.get("_gather_regions", {}).values():
        reg.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_gpu_result_equals_one_gpu_result():
    import torch.multiprocessing as mp

    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    import queue as _queue
    import time

    report, t0 = None, time.time()
    while report is None and time.time() - t0 < 420:
        try:
            report = q.get(timeout=2)
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    if report is None or "error" in report:
        for p in procs:
            if p.is_alive():
                p.kill()
        raise AssertionError(f"multi-GPU workers failed: {report}")
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert report and all(report.values()), report
