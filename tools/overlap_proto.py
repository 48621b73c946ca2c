"""Prototype: sweep of sub-batch k+1 on one stream while sub-batch k is refined / filtered / verified on another.
Prints step times for the bench shape (1,225 pairs) and a 2,048-pair batch, serial vs overlapped, for 1 / 2 / 4 / 8 sub-batches."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sfm-project_b200")]
import numpy as np
import torch

import sfm_b200
from sfm_b200 import _lib, synth
from sfm_b200.matcher import filter_params
from sfm_b200.ransac import ransac_params

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sc = synth.make_scene(n_img, 8192, seed=2001)
pairs_all = synth.exhaustive_pairs(n_img)
bank = sfm_b200.DescriptorBank(n_img, 8192)
bank.put(0, sc.desc, xy=sc.xy)
dev = bank.device
L = _lib.lib()
cap = bank.feat_stride
fprm = filter_params(0.75, "cv2_f32", False)
rprm = ransac_params(thr=3.0, confidence=0.99, max_iters=2000, solver="8pt", score="sym_epipolar", lo=False, seed=1, min_inliers=0)
mprm = _lib.MatchParams()
mprm.sweep_only = 4
mprm.prefilter_mode, mprm.prefilter_ratio = fprm.ratio_mode, fprm.ratio
mprm.prefilter_num, mprm.prefilter_den = int(fprm.ratio_num), int(fprm.ratio_den)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


class Bufs:
    def __init__(self, B):
        i32 = dict(dtype=torch.int32, device=dev)
        self.knn = torch.empty((B, cap, 4), **i32)
        self.blk = torch.empty(B * (cap // 256), **i32)
        self.counts, self.offsets = torch.zeros(B, **i32), torch.zeros(B + 1, **i32)
        self.matches = torch.empty((B * cap, 3), **i32)
        self.corr = torch.empty((B * cap, 4), dtype=torch.float32, device=dev)
        self.mask = torch.empty(B * cap, dtype=torch.uint8, device=dev)
        self.F = torch.zeros((B, 9), dtype=torch.float64, device=dev)
        self.ninl, self.iters = torch.zeros(B, **i32), torch.zeros(B, **i32)


def sweep(pairs_d, b, st):
    _lib.check(L.sfm_match_knn2(bank.handle, _lib.ptr(pairs_d), pairs_d.shape[0], C.byref(mprm), _lib.ptr(b.knn), st), "sweep")


def finish(pairs_d, ids_d, b, st):
    P = pairs_d.shape[0]
    _lib.check(L.sfm_refine_filter_packed(bank.handle, _lib.ptr(pairs_d), P, C.byref(fprm), None, _lib.ptr(b.knn), _lib.ptr(b.blk), _lib.ptr(b.counts),
                                          _lib.ptr(b.offsets), _lib.ptr(b.matches), _lib.ptr(b.corr), st), "finish")
    _lib.check(L.sfm_ransac_f_packed(_lib.ptr(b.corr), _lib.ptr(b.offsets), P, cap, _lib.ptr(ids_d), None, C.byref(rprm), _lib.ptr(b.F), _lib.ptr(b.ninl),
                                     _lib.ptr(b.mask), _lib.ptr(b.iters), st), "ransac")


def run(P, k, overlap):
    pairs = pairs_all[:P]
    cuts = np.linspace(0, P, k + 1).round().astype(int)
    B = int(np.diff(cuts).max())
    bufs = [Bufs(B), Bufs(B)]
    pd = torch.from_numpy(pairs).to(dev)
    ids = torch.arange(P, dtype=torch.int32, device=dev)
    main = torch.cuda.current_stream(dev)
    post = torch.cuda.Stream(device=dev) if overlap else main
    ev_sw = [torch.cuda.Event() for _ in range(k)]
    ev_fin = [torch.cuda.Event() for _ in range(k)]

    def step():
        for i in range(k):
            a, b_ = cuts[i], cuts[i + 1]
            bb = bufs[i % 2]
            if overlap and i >= 2:
                main.wait_event(ev_fin[i - 2])                       # the scratch of two sub-batches ago has been consumed
            sweep(pd[a:b_], bb, C.c_void_p(main.cuda_stream))
            if overlap:
                ev_sw[i].record(main)
                post.wait_event(ev_sw[i])
            finish(pd[a:b_], ids[a:b_], bb, C.c_void_p(post.cuda_stream))
            if overlap:
                ev_fin[i].record(post)
        if overlap:
            main.wait_stream(post)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    res = (bufs[(k - 1) % 2].ninl[: cuts[k] - cuts[k - 1]].sum().item(), bufs[(k - 1) % 2].counts[: cuts[k] - cuts[k - 1]].sum().item())
    return float(np.median(ms)), res


for P in (1225, min(2048, len(pairs_all))):
    for k in (1, 2, 4, 8):
        s_ms, r1 = run(P, k, False)
        o_ms, r2 = run(P, k, True)
        print(f"P={P} sub-batches={k}: serial {s_ms:.3f} ms   overlapped {o_ms:.3f} ms   ({'same' if r1 == r2 else 'DIFFERENT'} last-batch sums {r1})", flush=True)
