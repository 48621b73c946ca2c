"""Descriptor bank: descriptors packed once, resident in HBM (C ABI: sfm_bank_*).

Replaces the implicit hand-off between ``orb.detectAndCompute`` and ``bf.match`` in the reference
(code/feature_matching.py:44-50): instead of re-extracting and re-uploading per pair, every image's
descriptors are packed once (offset int8 + K-extension + norms) and all pairs are
matched from the bank.  The storage is one torch uint8 tensor owned by the caller side, so a
multi-GPU run can ``torch.distributed.broadcast`` it as is.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

METRICS = {"l2": _lib.METRIC_L2, "hamming": _lib.METRIC_HAMMING}
DIMS = {"l2": 128, "hamming": 32}


class DescriptorBank:
    def __init__(self, max_images: int, max_feats: int, metric: str = "l2", device=None):
        if metric not in METRICS:
            raise ValueError(f"metric must be 'l2' or 'hamming', got {metric!r}")
        if not torch.cuda.is_available():
            raise _lib.SfmError("DescriptorBank needs a CUDA device (sm_100a); there is no CPU fallback")
        self.metric = metric
        self.dim = DIMS[metric]
        self.max_images, self.max_feats = int(max_images), int(max_feats)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        L = _lib.lib()
        nbytes = C.c_size_t(0)
        _lib.check(L.sfm_bank_storage_bytes(self.max_images, self.max_feats, METRICS[metric], C.byref(nbytes)), "sfm_bank_storage_bytes")
        # +1024 so the base can be aligned to 1024 B regardless of the allocator
        self._raw = torch.zeros(nbytes.value + 1024, dtype=torch.uint8, device=self.device)
        off = (-self._raw.data_ptr()) % 1024
        self.storage = self._raw[off: off + nbytes.value]
        handle = C.c_void_p()
        _lib.check(
            L.sfm_bank_create(self.device.index or 0, self.max_images, self.max_feats, METRICS[metric],
                              C.c_void_p(self.storage.data_ptr()), nbytes.value, C.byref(handle)),
            "sfm_bank_create",
        )
        self._h = handle
        lay = (C.c_int64 * 6)()
        _lib.check(L.sfm_bank_layout(self._h, lay), "sfm_bank_layout")
        self.feat_stride = int(lay[0])
        self._off = {"desc": int(lay[1]), "ext": int(lay[2]), "norm": int(lay[3]), "xy": int(lay[4]), "count": int(lay[5])}
        self.n_images = 0
        self._counts_host = np.zeros(self.max_images, np.int32)

    # ------------------------------------------------------------------ views into the storage
    def section(self, name: str) -> torch.Tensor:
        keys = list(self._off)
        i = keys.index(name)
        end = self._off[keys[i + 1]] if i + 1 < len(keys) else self.storage.numel()
        return self.storage[self._off[name]: end]

    @property
    def counts(self) -> torch.Tensor:
        return self.section("count")[: 4 * self.max_images].view(torch.int32)

    @property
    def xy(self) -> torch.Tensor:
        rows = self.max_images * self.feat_stride
        return self.section("xy")[: rows * 8].view(torch.float32).view(self.max_images, self.feat_stride, 2)

    @property
    def norms(self) -> torch.Tensor:
        rows = self.max_images * self.feat_stride
        return self.section("norm")[: rows * 4].view(torch.int32).view(self.max_images, self.feat_stride)

    @property
    def handle(self):
        if self._h is None:
            raise _lib.SfmError("bank was destroyed")
        return self._h

    # ------------------------------------------------------------------ filling
    def put(self, first_image: int, desc, counts=None, xy=None) -> None:
        """Pack ``desc`` [n_images, n, dim] uint8 (numpy = host, copied from pinned memory; torch cuda = in place)
        into images [first_image, first_image + n_images)."""
        desc_t = self._to_device(desc, torch.uint8)
        if desc_t.dim() == 2:
            desc_t = desc_t.unsqueeze(0)
        if desc_t.dim() != 3 or desc_t.shape[2] != self.dim:
            raise ValueError(f"descriptors must be [n_images, n, {self.dim}] uint8, got {tuple(desc_t.shape)}")
        n_img, n = int(desc_t.shape[0]), int(desc_t.shape[1])
        if n == 0 or n_img == 0:
            raise ValueError("empty descriptor batch; use counts=0 rows instead")
        if n > self.feat_stride:
            raise ValueError(f"{n} features exceed the bank's feat_stride {self.feat_stride}")
        counts_t = None
        if counts is not None:
            counts_np = np.asarray(counts, np.int32).reshape(n_img)
            counts_t = torch.from_numpy(counts_np).to(self.device)
        else:
            counts_np = np.full(n_img, n, np.int32)
        xy_t = None
        if xy is not None:
            xy_t = self._to_device(xy, torch.float32).reshape(n_img, n, 2).contiguous()
        _lib.check(
            _lib.lib().sfm_bank_put_batch(self.handle, int(first_image), n_img, _lib.ptr(desc_t), n, _lib.ptr(counts_t),
                                          _lib.ptr(xy_t), _lib.current_stream_ptr(self.device)),
            "sfm_bank_put_batch",
        )
        # keep sources alive until the pack kernel has consumed them
        self._keepalive = (desc_t, counts_t, xy_t)
        self._counts_host[first_image: first_image + n_img] = np.minimum(counts_np, n)
        self.n_images = max(self.n_images, first_image + n_img)

    def mark_filled(self, n_images: int, counts_host=None) -> None:
        """The storage was filled by a broadcast from another rank."""
        _lib.check(_lib.lib().sfm_bank_mark_filled(self.handle, int(n_images)), "sfm_bank_mark_filled")
        self.n_images = int(n_images)
        if counts_host is not None:
            self._counts_host[: n_images] = np.asarray(counts_host, np.int32)[: n_images]

    def counts_host(self) -> np.ndarray:
        return self._counts_host[: self.n_images].copy()

    def _to_device(self, a, dtype) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            t = a.to(dtype=dtype)
            if t.device.type == "cpu":
                t = t.contiguous()
                t = (t if t.is_pinned() else t.pin_memory()).to(self.device, non_blocking=True)
            elif t.device != self.device:
                t = t.to(self.device)
            return t.contiguous()
        arr = np.ascontiguousarray(a, dtype={torch.uint8: np.uint8, torch.float32: np.float32}[dtype])
        return torch.from_numpy(arr).pin_memory().to(self.device, non_blocking=True)

    def destroy(self) -> None:
        if getattr(self, "_h", None) is not None:
            _lib.lib().sfm_bank_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def build_bank(descriptor_arrays, keypoint_xy=None, metric: str = "l2", device=None, max_feats=None) -> DescriptorBank:
    """Bank from a list of per-image descriptor arrays ([n_i, dim] uint8, ragged; ``None``/empty allowed)
    and optional per-image keypoint coordinates ([n_i, 2])."""
    dim = DIMS[metric]
    n_img = len(descriptor_arrays)
    ns = [0 if d is None else int(len(d)) for d in descriptor_arrays]
    cap = max(max(ns, default=0), 1) if max_feats is None else int(max_feats)
    bank = DescriptorBank(n_img, cap, metric, device)
    dense = np.zeros((n_img, cap, dim), np.uint8)
    xy = np.zeros((n_img, cap, 2), np.float32)
    for k, d in enumerate(descriptor_arrays):
        if ns[k]:
            dense[k, : ns[k]] = np.asarray(d, np.uint8).reshape(ns[k], dim)
            if keypoint_xy is not None and keypoint_xy[k] is not None:
                xy[k, : ns[k]] = np.asarray(keypoint_xy[k], np.float32).reshape(ns[k], 2)
    bank.put(0, dense, counts=np.asarray(ns, np.int32), xy=xy)
    return bank
