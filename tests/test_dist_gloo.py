"""world_size-2 / -3 gloo tests of the pair-sharding plumbing (CPU tensors; the same code runs under NCCL).

The reference collects every pair's matches in one list (code/pipeline.py:43-47); sharded, that list is assembled on
rank 0 by ``dist.gather_summaries`` (fixed-size per-pair summaries) and ``dist.GatherRegion`` (packed rows).  These tests
push hand-made per-pair payloads that encode the global pair index through both and check the assembled result."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_every_pair_once():
    from sfm_b200.dist import layout, partition

    for n in (0, 1, 7, 1225, 19900):
        for ws in (1, 2, 4, 8):
            for mode in ("block", "cyclic"):
                got = np.concatenate([partition(n, r, ws, mode) for r in range(ws)])
                assert sorted(got.tolist()) == list(range(n))
                lay = layout(n, ws, mode)
                assert sorted(lay.perm.tolist()) == list(range(n)) and len(lay.src_rows) == n
    b = partition(10, 1, 4, "block")
    assert b.tolist() == [3, 4, 5]
    assert partition(9, 3, 4, "block").tolist() == []                  # ceil(9/4) * 3 >= 9: the last rank is empty
    with pytest.raises(ValueError):
        partition(10, 0, 2, "zigzag")


def _fake_result(mine, with_h=False, with_pose=False):
    """Per-pair summaries that encode the global pair index; n_matches = 3 + (index % 5) rows per pair."""
    n = len(mine)
    i64 = torch.as_tensor(np.asarray(mine, np.int64))
    F = (i64.to(torch.float64).view(n, 1, 1) + torch.arange(9, dtype=torch.float64).view(1, 3, 3) / 16.0) if n else torch.zeros((0, 3, 3), dtype=torch.float64)
    r = SimpleNamespace(n_matches=(3 + i64 % 5).to(torch.int32), n_inliers=(i64 % 3).to(torch.int32), iters=torch.full((n,), 7, dtype=torch.int32), F=F)
    if with_h:
        r.H, r.n_inliers_h = F + 0.5, (i64 % 2).to(torch.int32)
    if with_pose:
        r.R, r.t, r.n_pose = F * 2.0, torch.stack([i64.to(torch.float64), torch.ones(n, dtype=torch.float64), torch.full((n,), 2.0, dtype=torch.float64)], 1), (i64 % 4).to(torch.int32)
    return r


def _rows_of(pair_index, n_rows):
    """Packed rows of one pair: matches[k] = (pair, k, 1000 * pair + k), inlier[k] = (pair + k) & 1."""
    k = np.arange(n_rows)
    return np.stack([np.full(n_rows, pair_index), k, 1000 * pair_index + k], 1).astype(np.int32), ((pair_index + k) & 1).astype(np.uint8)


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from sfm_b200.dist import GatherRegion, gather_summaries, layout

    ok = True
    cpu = torch.device("cpu")
    for n_total, mode in ((11, "block"), (11, "cyclic"), (ws - 1, "block"), (0, "block")):
        lay = layout(n_total, ws, mode)
        mine = lay.owned[rank]
        # ---- summaries: plain, and with the optional stages requested (a rank with an EMPTY block must still send the wide tile)
        summ = gather_summaries(_fake_result(mine), mine, n_total, 0, mode=mode)
        res_x = _fake_result(mine, True, True)
        if len(mine) == 0:                                           # what match_and_verify returns for P == 0 without the fix: no H / R
            res_x = _fake_result(mine)
        summ_x = gather_summaries(res_x, mine, n_total, 0, mode=mode, homography=True, pose=True)
        # ---- packed rows through the sendrecv transport
        caps = [n * 8 for n in lay.sizes]
        reg = GatherRegion(caps, ("matches", "inlier"), 0, cpu, "sendrecv")
        rows = 0
        for p in mine:
            m, k = _rows_of(int(p), 3 + int(p) % 5)
            reg.local["matches"][rows: rows + len(m)] = torch.from_numpy(m)
            reg.local["inlier"][rows: rows + len(m)] = torch.from_numpy(k)
            rows += len(m)
        reg.exchange(rows, summ["n_matches"] if rank == 0 else None, lay)
        if rank == 0:
            want = _fake_result(np.arange(n_total), True, True)
            for name in ("n_matches", "n_inliers", "iters", "F"):
                ok &= torch.equal(summ[name], getattr(want, name)) and torch.equal(summ_x[name], getattr(want, name))
            for name in ("H", "n_inliers_h", "R", "t", "n_pose"):
                ok &= torch.equal(summ_x[name], getattr(want, name)) and name not in summ
            ok &= summ_x["F"].shape == (n_total, 3, 3) and summ_x["t"].shape == (n_total, 3)
            start = reg.row_starts(summ["n_matches"], lay)
            for p in range(n_total):
                m, k = _rows_of(p, 3 + p % 5)
                s = int(start[p])
                ok &= np.array_equal(reg.full["matches"][s: s + len(m)].numpy(), m) and np.array_equal(reg.full["inlier"][s: s + len(m)].numpy(), k)
            # rank r's rows sit inside its slice, in its processing order
            for r in range(ws):
                if lay.sizes[r]:
                    ok &= int(start[lay.owned[r][0]]) == int(reg.base[r])
        else:
            assert summ is None and summ_x is None
        # a rank that sends more rows than its slice holds is caught on the gathering rank
        reg.close()
    with pytest.raises(ValueError):
        gather_summaries(_fake_result(np.arange(3)), np.arange(3), 11, 0)            # not this rank's partition
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 3])
def test_gather_gloo(ws):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + ws
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
