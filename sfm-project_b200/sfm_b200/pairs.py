"""Pair-list producers and the reference's pair bookkeeping.

``code/pipeline.py:38-40`` enumerates the ORDERED pairs (i, j), i != j, inside its double loop; BASELINE.json's
configurations use unordered exhaustive lists and a sequential window.  ``Pair`` mirrors the
reference's record (``code/pipeline.py:6-9``; filled at ``:43-47``) and ``to_reference_pairs`` rebuilds the list the
reference's loop produces from a batched result.
"""
from __future__ import annotations

import numpy as np

from .synth import exhaustive_pairs, ordered_pairs, windowed_pairs  # noqa: F401  (re-exported: the pair-list producers)


def blocked_exhaustive_pairs(n_images: int, block: int = 32) -> np.ndarray:
    """The unordered exhaustive pair list (the same N(N-1)/2 pairs as ``exhaustive_pairs``) ordered by ``block`` x ``block``
    squares of the (i, j) triangle, (i, j)-sorted inside a square.  The (i, j)-sorted list of code/pipeline.py:38-40 walks all
    N train images for every query image -- once the bank outgrows the 126 MB L2 (200 images x 8192 features = 264 MB) every
    train tile then comes from HBM; in this order the pairs in flight touch 2 x ``block`` images (~86 MB at 32).  Results
    follow the order of the list the caller passes, whichever it is."""
    i, j = np.triu_indices(n_images, k=1)
    key = (i // block).astype(np.int64) * ((n_images + block - 1) // block) + (j // block)
    order = np.argsort(key, kind="stable")
    return np.stack([i[order], j[order]], axis=1).astype(np.int32)


class Pair:
    """Same attributes as the reference's ``Pair`` (code/pipeline.py:6-9)."""

    img_inx_1 = -1
    img_inx_2 = -1
    matches = []

    def __repr__(self):
        return f"Pair({self.img_inx_1}, {self.img_inx_2}, {len(self.matches)} matches)"


def to_reference_pairs(host: dict, *, inliers_only: bool = False, min_matches: int = 1, as_dmatch: bool = True) -> list:
    """The ``pair_matches`` list of ``code/pipeline.py:36-49`` from ``VerifiedPairs.to_host()``: one ``Pair`` per image pair
    that has matches (``if match:``, code/pipeline.py:42), in pair-list order, ``matches`` a ``list[cv2.DMatch]`` with cv2's
    float32 L2 distances (or int32 rows ``(queryIdx, trainIdx, squared distance)`` when ``as_dmatch`` is False).
    ``inliers_only`` keeps the geometrically verified matches (the stage the reference left empty, code/pipeline.py:60-65);
    ``min_matches`` drops weak pairs (scene-graph pruning)."""
    out = []
    off, rows, inl = host["offsets"], host["matches"], host["inlier"]
    if as_dmatch:
        import cv2
    for p, (i, j) in enumerate(np.asarray(host["pairs"]).reshape(-1, 2)):
        r = rows[off[p]: off[p + 1]]
        if inliers_only:
            r = r[inl[off[p]: off[p + 1]].astype(bool)]
        if len(r) < max(min_matches, 1):
            continue
        pair = Pair()
        pair.img_inx_1, pair.img_inx_2 = int(i), int(j)
        if as_dmatch:
            d = np.sqrt(r[:, 2].astype(np.float32))
            pair.matches = [cv2.DMatch(int(a), int(b), 0, float(c)) for a, b, c in zip(r[:, 0], r[:, 1], d)]
        else:
            pair.matches = r.copy()
        out.append(pair)
    return out
