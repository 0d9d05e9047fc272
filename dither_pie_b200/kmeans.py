"""K-means palette extraction: ColorReducer.generate_kmeans_palette (dithering_lib.py:1845-1857).

The reference calls ``sklearn.cluster.KMeans(n_clusters, random_state).fit`` on at most 10 000
sampled pixels.  Here the seeding (k-means++, a K-step sequential sampling procedure that must
consume numpy's RandomState exactly like sklearn does) is host set-up, and the Lloyd iterations
-- the per-pixel work -- run on the GPU with exact integer centroid sums:

    dp_kmeans_accumulate   assignment + per-cluster (sum r, sum g, sum b, count) as u64
    [all-reduce of the K x 4 integers across ranks when pixels are sharded over GPUs]
    dp_kmeans_update       centres = sums / count, squared centre shift (sklearn's stop value)

Because the sums are integers the centres are identical for any number of shards.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import List, Optional, Tuple

import numpy as np

from . import _capi
from ._capi import DeviceBuffer, check, lib

MAX_ITER = 300      # sklearn default
TOL = 1e-4          # sklearn default (relative to the mean per-feature variance)
SAMPLE = 10000      # dithering_lib.py:1850


def kmeans_plusplus(Xc: np.ndarray, k: int, seed) -> np.ndarray:
    """sklearn's greedy k-means++ (``_kmeans_plusplus``, n_local_trials = 2 + int(log k)) on the
    centred f64 data, drawing from RandomState(seed) in the same order."""
    rs = seed if isinstance(seed, np.random.RandomState) else np.random.RandomState(seed)
    n = Xc.shape[0]
    sq = np.einsum("ij,ij->i", Xc, Xc)
    trials = 2 + int(np.log(k))
    centers = np.empty((k, Xc.shape[1]), Xc.dtype)

    def dist_to(c):
        d = sq - 2.0 * (Xc @ c.T).T + np.einsum("ij,ij->i", c, c)[:, None]
        np.maximum(d, 0, out=d)
        return d

    first = rs.choice(n, p=np.full(n, 1.0 / n))
    centers[0] = Xc[first]
    closest = dist_to(centers[0:1])[0]
    pot = closest.sum()
    for c in range(1, k):
        draws = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(closest, dtype=np.float64), draws)
        np.clip(cand, None, n - 1, out=cand)
        dc = dist_to(Xc[cand])
        np.minimum(closest, dc, out=dc)
        pots = dc.sum(axis=1)
        b = int(np.argmin(pots))
        pot, closest = pots[b], dc[b]
        centers[c] = Xc[cand[b]]
    return centers


class LloydResult(tuple):
    """(centres, n_iter) with the extra device-side diagnostics as attributes."""

    def __new__(cls, centers, n_iter, shift2, ties, empty_iters):
        self = super().__new__(cls, (centers, n_iter))
        self.shift2, self.ties, self.empty_iters = shift2, ties, empty_iters
        return self


def lloyd_device(pixels_ptr: int, n: int, centers: np.ndarray, tol: float,
                 max_iter: int = MAX_ITER, comm: Optional[int] = None, stream=None,
                 check_every: int = 4, p2p=None) -> Tuple[np.ndarray, int]:
    """Lloyd iterations on device-resident u8 pixels [n,3] through ``dp_kmeans_lloyd``: the whole
    loop runs on the stream with the stop test on the device; the host looks at the flag every
    ``check_every`` iterations.  ``centers`` f64 [K,3] (uncentred).  When the pixels are one shard
    of a multi-GPU job pass ``comm`` (a dp_nccl_comm_create handle): the K x 4 integer sums are
    all-reduced with ncclAllReduce on the same stream, so every rank ends with identical centres;
    or ``p2p`` (from ``distributed.p2p_exchange()``): the kernels push the sums into the peers'
    inboxes over NVLink themselves, no NCCL launch.
    Returns (centres, n_iter); the result also carries ``.ties`` (samples exactly equidistant from
    their two nearest centres, summed over the iterations), ``.shift2`` and ``.empty_iters``."""
    K = int(centers.shape[0])
    c_host = np.ascontiguousarray(centers, np.float64).copy()
    n_iter, shift2 = C.c_int(0), C.c_double(0.0)
    ties, empty = C.c_ulonglong(0), C.c_int(0)
    if p2p is not None:
        # (rank, world, inbox pointers, epoch): the kernels exchange the sums over peer memory
        rank, world, inboxes, epoch = p2p
        arr = (C.c_void_p * world)(*[C.c_void_p(int(x)) for x in inboxes])
        check(lib().dp_kmeans_lloyd_p2p(pixels_ptr, int(n), c_host.ctypes.data, K, float(tol), int(max_iter),
                                        int(rank), int(world), arr, int(epoch), int(check_every),
                                        C.byref(n_iter), C.byref(shift2), C.byref(ties), C.byref(empty),
                                        stream), "dp_kmeans_lloyd_p2p")
    else:
        check(lib().dp_kmeans_lloyd(pixels_ptr, int(n), c_host.ctypes.data, K, float(tol), int(max_iter),
                                    comm, int(check_every), C.byref(n_iter), C.byref(shift2),
                                    C.byref(ties), C.byref(empty), stream), "dp_kmeans_lloyd")
    return LloydResult(c_host, int(n_iter.value), float(shift2.value), int(ties.value),
                       int(empty.value))


def kmeans_fit(sample_u8: np.ndarray, k: int, random_state=42) -> Tuple[np.ndarray, int]:
    """Pre-truncation centres (f64 [k,3]) and iteration count for the given pixels."""
    _capi.ensure_device()
    pix = np.ascontiguousarray(sample_u8, np.uint8).reshape(-1, 3)
    X = pix.astype(np.float64)
    mean = X.mean(axis=0)
    Xc = X - mean
    tol = float(np.mean(np.var(Xc, axis=0)) * TOL)
    init = kmeans_plusplus(Xc, k, random_state) + mean
    buf = DeviceBuffer(max(pix.nbytes, 4)).upload(pix)
    try:
        return lloyd_device(buf.ptr, pix.shape[0], init, tol)
    finally:
        buf.free()


def kmeans_palette(arr_u8: np.ndarray, num_colors: int, random_state=42) -> List[tuple]:
    """generate_kmeans_palette on a uint8 image: the 10 000-pixel sub-sample uses the global
    ``random`` module exactly like the reference (:1850-1853; seed it to reproduce), centres are
    truncated with ``astype(int)`` (:1856)."""
    pix = np.asarray(arr_u8, np.uint8).reshape(-1, 3)
    if len(pix) > SAMPLE:
        pix = pix[random.sample(range(len(pix)), SAMPLE)]
    res = kmeans_fit(pix, num_colors, random_state)
    if getattr(res, "empty_iters", 0):
        # sklearn relocates an empty cluster to the sample farthest from its centre
        # (_relocate_empty_clusters_dense, an argpartition whose order is numpy's); here an empty
        # cluster keeps its centre.  Only images with fewer distinct colours than clusters get here.
        import warnings
        warnings.warn("k-means: an empty cluster occurred (fewer distinct colours than clusters?); "
                      "the palette can differ from scikit-learn's, which relocates empty clusters",
                      RuntimeWarning, stacklevel=2)
    return [tuple(c) for c in res[0].astype(int)]
