"""GPU bring-up diagnostics (prints, does not assert).  Usage: python tools/bringup.py <stage> ...
Each stage runs in its own process under `timeout` (tools/bringup.sh) so a trap in one kernel
cannot poison the others."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sfm-project_b200"))

import numpy as np
import torch

import sfm_b200
from sfm_b200 import synth, matcher
from oracle import match_oracle as mo


def small_bank(n_feats=(700, 900), seed=0):
    rng = np.random.default_rng(seed)
    B = synth.sift_like(rng, n_feats[1])
    A = synth.sift_like(rng, n_feats[0])
    k = min(n_feats) // 2
    A[:k] = synth.observe(rng, B[rng.permutation(n_feats[1])[:k]])
    bank = sfm_b200.build_bank([A, B])
    return A, B, bank


def expected_acc(A, B):
    a = A.astype(np.int64) - 128
    b = B.astype(np.int64) - 128
    dot = a @ b.T
    h = (b * b).sum(1) >> 1
    return dot, (1 << 20) - h


def stage_pack():
    A, B, bank = small_bank()
    torch.cuda.synchronize()
    nb = ((B.astype(np.int64) - 128) ** 2).sum(1)
    got = bank.norms[1, : len(B)].cpu().numpy()
    print("norm ok:", np.array_equal(got, nb), "counts:", bank.counts.cpu().numpy()[:2])
    ext = bank.section("ext").cpu().numpy()
    fs = bank.feat_stride
    tile0 = ext[(fs // 128) * 4096: (fs // 128) * 4096 + 4096].reshape(2, 128, 16)
    e = np.concatenate([tile0[0], tile0[1]], axis=1).astype(np.int64)   # [128, 32]
    w = np.array([255] * 24 + [1] + [0] * 7)
    v = (e * w).sum(1)
    print("ext ok:", np.array_equal(v, (1 << 20) - (nb[:128] >> 1)))


def stage_simt():
    A, B, bank = small_bank()
    knn = sfm_b200.knn2(bank, [[0, 1], [1, 0]], impl="simt")
    torch.cuda.synchronize()
    knn = knn.cpu().numpy()
    for p, (X, Y) in enumerate([(A, B), (B, A)]):
        i1, d1, i2, d2 = mo.l2_knn2(X, Y)
        g = knn[p, : len(X)]
        print(f"simt pair {p}: idx1 {np.array_equal(g[:,0], i1)} d1 {np.array_equal(g[:,1], d1)} idx2 {np.array_equal(g[:,2], i2)} d2 {np.array_equal(g[:,3], d2)}"
              f" pad rows -1: {(knn[p, len(X):] == -1).all()}")


def stage_tile(mode):
    A, B, bank = small_bank()
    acc, knn = matcher.debug_tc_tile(bank, [0, 1], mode)
    torch.cuda.synchronize()
    acc = acc.cpu().numpy().astype(np.int64)
    dot, ext = expected_acc(A[:256], B[:128])
    want = {0: dot + ext[None, :], 1: dot, 2: np.broadcast_to(ext[None, :], dot.shape)}[mode]
    bad = acc != want
    print(f"tile mode {mode}: mismatches {bad.sum()} / {bad.size}")
    if bad.any():
        r, c = np.nonzero(bad)
        print("  first bad (row,col,got,want):", [(int(a), int(b), int(acc[a, b]), int(want[a, b])) for a, b in zip(r[:6], c[:6])])
        print("  bad rows:", np.unique(r)[:20], " bad cols:", np.unique(c)[:20])
        print("  got[0,:8]", acc[0, :8], "want[0,:8]", want[0, :8])
    if mode == 0:
        knn = knn.cpu().numpy()
        i1, d1, i2, d2 = mo.l2_knn2(A, B)
        g = knn[: len(A)]
        print(f"  1-CTA full pair knn: idx1 {np.array_equal(g[:,0], i1)} d1 {np.array_equal(g[:,1], d1)} idx2 {np.array_equal(g[:,2], i2)} d2 {np.array_equal(g[:,3], d2)}")


def stage_tc():
    for n_feats, seed in (((700, 900), 0), ((2048, 2048), 1), ((8192, 8192), 2), ((300, 5000), 3), ((1, 2), 4)):
        A, B, bank = small_bank(n_feats, seed)
        pairs = [[0, 1], [1, 0], [0, 0]]
        knn_t = sfm_b200.knn2(bank, pairs, impl="tcgen05")
        knn_s = sfm_b200.knn2(bank, pairs, impl="simt")
        torch.cuda.synchronize()
        same = torch.equal(knn_t, knn_s)
        print(f"tc vs simt n={n_feats}: identical {same}")
        if not same:
            d = (knn_t != knn_s).any(2).cpu().numpy()
            for p in range(3):
                rows = np.nonzero(d[p])[0]
                print(f"   pair {p}: {len(rows)} differing rows, first {rows[:8]}")
                for r in rows[:3]:
                    print("     tc", knn_t[p, r].tolist(), "simt", knn_s[p, r].tolist())
        if max(n_feats) <= 2048:
            i1, d1, i2, d2 = mo.l2_knn2(A, B)
            g = knn_t[0, : len(A)].cpu().numpy()
            print(f"   vs oracle: {np.array_equal(g[:,0], i1) and np.array_equal(g[:,1], d1) and np.array_equal(g[:,2], i2) and np.array_equal(g[:,3], d2)}")


def stage_time():
    sc = synth.make_scene(8, 8192, seed=1)
    bank = sfm_b200.DescriptorBank(8, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = synth.exhaustive_pairs(8)
    pairs = np.concatenate([pairs] * 8)           # 224 pairs
    out = torch.empty((len(pairs), bank.feat_stride, 4), dtype=torch.int32, device="cuda")
    matcher.refine_stats(True)
    sfm_b200.knn2(bank, pairs, impl="tcgen05", out=out)
    torch.cuda.synchronize()
    br, cand = matcher.refine_stats(False)
    print(f"refine stats: {br} brute rows, {cand} candidates over {len(pairs)*8192} rows ({cand/(len(pairs)*8192):.1f}/row)")
    for impl in ("tcgen05", "sweep", "simt"):
        kw = dict(impl="tcgen05", sweep_only=True) if impl == "sweep" else dict(impl=impl)
        for _ in range(2):
            sfm_b200.knn2(bank, pairs, out=out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            sfm_b200.knn2(bank, pairs, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ops = len(pairs) * 2.0 * 8192 * 8192 * 128
        print(f"{impl}: {ms:.3f} ms for {len(pairs)} pairs -> {len(pairs)/ms*1e3:.0f} pairs/s, {ops/ms/1e9:.1f} TOP/s")
    ms, rate = sfm_b200.probe_int8_peak(0, 8192)
    print(f"probe int8 mma: {ms:.3f} ms, {rate/1e12:.1f} TOP/s (algorithmic, K-extension not credited)")
    a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    print(f"torch._int_mm 8192^3: {2*8192**3*10/e0.elapsed_time(e1)/1e9:.1f} TOP/s")


def stage_filter():
    A, B, bank = small_bank((1500, 1800), 5)
    for mode in ("cv2_f32", "exact_int", None):
        for mutual in (False, True):
            mb = sfm_b200.match_pairs(bank, [[0, 1]], ratio=0.75 if mode else None, ratio_mode=mode, mutual=mutual, impl="simt")
            q, t, d = mb.to_host()[0]
            oq, ot, od = mo.match_l2(A, B, ratio=0.75 if mode else None, ratio_mode=mode or "cv2_f32", mutual=mutual)
            print(f"filter mode={mode} mutual={mutual}: n={len(q)} ok={np.array_equal(q, oq) and np.array_equal(t, ot) and np.array_equal(d, od)}")


def stage_hamming():
    rng = np.random.default_rng(3)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    a = rng.integers(0, 256, (420, 32), dtype=np.uint8)
    a[:200] = b[rng.permutation(500)[:200]] ^ (rng.random((200, 32)) < 0.03).astype(np.uint8)
    a[7] = a[0]
    b[499] = b[0]
    bank = sfm_b200.build_bank([a, b], metric="hamming")
    mb = sfm_b200.match_pairs_hamming(bank, [[0, 1], [1, 0]], 26)
    res = mb.to_host()
    for p, (X, Y) in enumerate([(a, b), (b, a)]):
        oq, ot, od = mo.match_hamming_reference(X, Y, 26)
        q, t, d = res[p]
        print(f"hamming pair {p}: n={len(q)} (oracle {len(oq)}) ok={np.array_equal(q, oq) and np.array_equal(t, ot) and np.array_equal(d, od)}")


def stage_ransac():
    from oracle import ransac_oracle as ro
    for solver in ("7pt", "8pt"):
        for lo in (False, True):
            for score in ("sym_epipolar", "sampson"):
                cs, counts = [], []
                data = []
                for k, (n, outl) in enumerate([(500, 0.3), (2000, 0.5), (6, 0.0), (1200, 0.6)]):
                    p1, p2, gt, _ = synth.two_view_correspondences(n, outlier_frac=outl, seed=40 + k)
                    data.append((p1, p2, gt))
                    c = np.zeros((2048, 4), np.float32)
                    c[:n, :2], c[:n, 2:] = p1, p2
                    cs.append(c)
                    counts.append(n)
                corr = torch.from_numpy(np.stack(cs)).cuda()
                t0 = time.time()
                vb = sfm_b200.verify_corr(corr, torch.tensor(counts, dtype=torch.int32), thr=3.0, confidence=0.99, max_iters=1024,
                                          solver=solver, score=score, lo=lo, seed=9)
                torch.cuda.synchronize()
                dt = time.time() - t0
                F, ninl, mask, iters = vb.F.cpu().numpy(), vb.n_inliers.cpu().numpy(), vb.mask.cpu().numpy(), vb.iters.cpu().numpy()
                msgs = []
                for k, (p1, p2, gt) in enumerate(data):
                    oF, om, on, oi = ro.ransac_f(p1, p2, pair_id=k, solver=int(solver[0]), score=0 if score == "sym_epipolar" else 1, thr=3.0,
                                                 max_iters=1024, confidence=0.99, seed=9, lo=lo)
                    okF = (oF is None and ninl[k] == 0) or (oF is not None and np.array_equal(oF, F[k]))
                    msgs.append(f"[n={counts[k]} ninl {ninl[k]}/{on} it {iters[k]}/{oi} mask {np.array_equal(mask[k,:counts[k]], om)} F {okF}]")
                print(f"ransac {solver} lo={lo} {score} ({dt*1e3:.0f} ms):", " ".join(msgs))


def stage_trace():
    import ctypes as C
    from sfm_b200 import _lib
    sc = synth.make_scene(8, 8192, seed=1)
    bank = sfm_b200.DescriptorBank(8, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    for n_pairs in (1, 28):
        pairs = synth.exhaustive_pairs(8)[:n_pairs]
        pairs_t = torch.from_numpy(pairs).cuda()
        knn = torch.empty((n_pairs, bank.feat_stride, 4), dtype=torch.int32, device="cuda")
        acc = torch.zeros((256, 128), dtype=torch.int32, device="cuda")
        for _ in range(2):
            _lib.check(_lib.lib().sfm_debug_tc_tile(bank.handle, _lib.ptr(pairs_t), 4 | (n_pairs << 8), _lib.ptr(knn), _lib.ptr(acc),
                                                    _lib.current_stream_ptr(bank.device)))
        torch.cuda.synchronize()
        tr = acc.cpu().numpy().view(np.int64).reshape(-1)[: 4 * 64 * 8].reshape(4, 64, 8)
        t0 = tr[0, 0, 0]
        rel = lambda x: int(x - t0) if x else -1
        print(f"--- trace n_pairs={n_pairs} (cycles relative to MMA tile 0 start)")
        print("tile | MMA rb0: top b_full t_empty issued | MMA rb1: same | EPI rb0: top l0 pre3 l3 rel pf end | EPI rb1: same | TMA: top issue")
        for t in list(range(0, 6)) + list(range(40, 46)):
            m = [rel(x) for x in tr[0, t, :8]]
            e0 = [rel(x) for x in tr[1, t, :7]]
            e1 = [rel(x) for x in tr[2, t, :7]]
            a = [rel(x) for x in tr[3, t, :2]]
            print(t, "|", *m, "|", *e0, "|", *e1, "|", *a)
        d = np.diff(tr[0, 8:60, 0])
        print("MMA tile period (cycles): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))


def stage_bounds():
    """Which side bounds the sweep: time the debug kernel on the full grid with parts of the work switched off."""
    import ctypes as C
    from sfm_b200 import _lib
    sc = synth.make_scene(8, 8192, seed=1)
    bank = sfm_b200.DescriptorBank(8, 8192)
    bank.put(0, sc.desc, xy=sc.xy)
    pairs = np.concatenate([synth.exhaustive_pairs(8)] * 8)           # 224 pairs
    n_pairs = len(pairs)
    pairs_t = torch.from_numpy(pairs).cuda()
    knn = torch.empty((n_pairs, bank.feat_stride, 4), dtype=torch.int32, device="cuda")
    acc = torch.zeros((256, 128), dtype=torch.int32, device="cuda")
    names = {5: "K-extension MMA only + production epilogue (epilogue bound)",
             6: "full MMA, epilogue releases at once (MMA + handshake bound)", 7: "full MMA, epilogue reads TMEM, no arithmetic"}
    for mode in (5, 6, 7):
        def run():
            _lib.check(_lib.lib().sfm_debug_tc_tile(bank.handle, _lib.ptr(pairs_t), mode | (n_pairs << 8), _lib.ptr(knn), _lib.ptr(acc),
                                                    _lib.current_stream_ptr(bank.device)))
        run(); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"mode {mode}: {ms:.3f} ms for {n_pairs} pairs = {ms*1e-3*1.965e9*148/(n_pairs*32*64):.0f} cycles per B tile @1965 MHz  -- {names[mode]}")


STAGES = {"bounds": stage_bounds, "pack": stage_pack, "simt": stage_simt, "tile1": lambda: stage_tile(1), "tile2": lambda: stage_tile(2),
          "tile0": lambda: stage_tile(0), "tc": stage_tc, "time": stage_time, "filter": stage_filter,
          "hamming": stage_hamming, "ransac": stage_ransac, "trace": stage_trace}

if __name__ == "__main__":
    for s in sys.argv[1:]:
        print(f"=== stage {s}", flush=True)
        t0 = time.time()
        try:
            STAGES[s]()
        except Exception as e:  # noqa: BLE001
            print(f"stage {s} raised {type(e).__name__}: {e}")
        print(f"=== stage {s} done in {time.time()-t0:.1f}s", flush=True)
