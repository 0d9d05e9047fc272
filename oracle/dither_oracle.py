"""CPU oracle for dither_pie's per-pixel hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package (dither_pie_b200/) never does; it fails
loudly when its CUDA library is missing.

This is a restatement ("port") of the reference algorithm, function by function, in
numpy + scipy for the vectorisable parts and in C (oracle/c/dp_oracle.c) for the serial
parts.  All ``file:line`` citations are into the reference repository
(dobrosketchkun/dither_pie), file ``dithering_lib.py`` unless another file is named.

Third-party code that owns part of the arithmetic (not vendored by the reference, no version
pins in the reference; versions below are the ones the golden vectors were generated with):
  scipy.spatial.KDTree 1.18.1   nearest / second-nearest palette entries (:339, :358, :554,
                                :748, :1229, :1612)
  numba 0.65.0                  JIT of _error_diffusion_numba (:212-308)
  scikit-learn 1.9.0            KMeans (:1854-1856)
  Pillow 12.2.0                 Image.resize(NEAREST) (video_processor.py:576)
  numpy 2.3.5                   NEP-50 promotion at every f32/f64 boundary

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build container by
tools/make_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
from scipy.spatial import KDTree, cKDTree

# ----------------------------------------------------------------------------------------
# C library (serial kernels)
# ----------------------------------------------------------------------------------------

_LIB = None


class _KdTreeC(ctypes.Structure):
    _fields_ = [
        ("n_nodes", ctypes.c_int32),
        ("n_points", ctypes.c_int32),
        ("split_dim", ctypes.c_void_p),
        ("split", ctypes.c_void_p),
        ("start_idx", ctypes.c_void_p),
        ("end_idx", ctypes.c_void_p),
        ("lesser", ctypes.c_void_p),
        ("greater", ctypes.c_void_p),
        ("indices", ctypes.c_void_p),
        ("data", ctypes.c_void_p),
        ("mins", ctypes.c_void_p),
        ("maxes", ctypes.c_void_p),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        from . import build as _build

        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_kdtree_query.restype = ctypes.c_int
        _LIB.orc_error_diffusion.restype = ctypes.c_int
        _LIB.orc_ostromoukhov.restype = ctypes.c_int
        _LIB.orc_hybrid.restype = ctypes.c_int
        _LIB.orc_perceptual.restype = ctypes.c_int
        _LIB.orc_adaptive.restype = ctypes.c_int
    return _LIB


# ----------------------------------------------------------------------------------------
# KD-tree export: the tree is built by scipy itself, flattened in pre-order
# ----------------------------------------------------------------------------------------

def export_kdtree(palette: np.ndarray) -> Dict[str, np.ndarray]:
    """Flatten scipy's KDTree(palette) (default leafsize=10) into plain arrays."""
    tree = KDTree(np.asarray(palette))
    root = cKDTree.tree.__get__(tree)
    split_dim, split, start, end, lesser, greater = [], [], [], [], [], []

    def visit(node) -> int:
        me = len(split_dim)
        split_dim.append(int(node.split_dim))
        split.append(float(node.split))
        start.append(int(node.start_idx))
        end.append(int(node.end_idx))
        lesser.append(-1)
        greater.append(-1)
        if node.split_dim != -1:
            lesser[me] = visit(node.lesser)
            greater[me] = visit(node.greater)
        return me

    visit(root)
    return {
        "split_dim": np.asarray(split_dim, np.int32),
        "split": np.asarray(split, np.float64),
        "start_idx": np.asarray(start, np.int32),
        "end_idx": np.asarray(end, np.int32),
        "lesser": np.asarray(lesser, np.int32),
        "greater": np.asarray(greater, np.int32),
        "indices": np.ascontiguousarray(tree.indices, np.int32),
        "data": np.ascontiguousarray(tree.data, np.float64),
        "mins": np.ascontiguousarray(tree.mins, np.float64),
        "maxes": np.ascontiguousarray(tree.maxes, np.float64),
    }


def _kd_struct(t: Dict[str, np.ndarray]) -> _KdTreeC:
    s = _KdTreeC()
    s.n_nodes = len(t["split_dim"])
    s.n_points = t["data"].shape[0]
    for name in ("split_dim", "split", "start_idx", "end_idx", "lesser", "greater",
                 "indices", "data", "mins", "maxes"):
        setattr(s, name, t[name].ctypes.data)
    return s


def kdtree_query_c(tree: Dict[str, np.ndarray], points: np.ndarray, k: int
                   ) -> Tuple[np.ndarray, np.ndarray]:
    """The C restatement of scipy's query; returns (squared distances f64 [n,k], idx [n,k])."""
    pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    n = pts.shape[0]
    idx = np.empty((n, k), np.int32)
    d2 = np.empty((n, k), np.float64)
    s = _kd_struct(tree)
    rc = _lib().orc_kdtree_query(ctypes.byref(s), ctypes.c_void_p(pts.ctypes.data),
                                 ctypes.c_int64(n), ctypes.c_int(k),
                                 ctypes.c_void_p(idx.ctypes.data), ctypes.c_void_p(d2.ctypes.data))
    if rc != 0:
        raise RuntimeError("orc_kdtree_query failed")
    return d2, idx


# ----------------------------------------------------------------------------------------
# Threshold sources
# ----------------------------------------------------------------------------------------

def _bayer_index(n: int) -> np.ndarray:
    """Classic recursive Bayer index matrix of side n (power of two)."""
    m = np.zeros((1, 1), np.int64)
    while m.shape[0] < n:
        m = np.block([[4 * m, 4 * m + 2], [4 * m + 3, 4 * m + 1]])
    return m


def bayer_matrix(size: str) -> np.ndarray:
    """Threshold tables of DitherUtils (:1705-1768), rebuilt procedurally.

    The reference tables are (index+1)/n^2 for 2x2, 8x8 and 16x16 and (index+0.5)/16 for 4x4.
    Its 8x8 table departs from that rule in two ways that are kept on purpose (parity):
    the right 4x4 block of rows 4..7 repeats the 4x4 table, and entries [3,6:8] are
    0.84375, 0.34375.  Its 16x16 table is NOT the canonical recursion: it is tiled from 8x8
    style blocks with sub-offsets; tests pin it against golden data extracted from the
    reference.  Unknown sizes fall back to 4x4 (:441-442).
    """
    if size == "2x2":
        return ((_bayer_index(2) + 1) / 4.0).astype(np.float32)
    if size in ("psx4x4", "psx"):
        psx = np.array([[1, 9, 3, 11], [13, 5, 15, 7], [3, 11, 1, 9], [15, 7, 13, 5]])
        return (psx / 16.0).astype(np.float32)
    b4 = ((_bayer_index(4) + 0.5) / 16.0)
    if size == "8x8":
        b8 = (_bayer_index(8) + 1) / 64.0
        b8[4:8, 4:8] = b4
        b8[3, 6] = 0.84375
        b8[3, 7] = 0.34375
        return b8.astype(np.float32)
    if size == "16x16":
        return _bayer16()
    return b4.astype(np.float32)


def _bayer16() -> np.ndarray:
    """The reference's 16x16 table (:1728-1761).

    Rows 0..7: columns 0..7 are (B8+1)/256-style values built from the canonical 16x16
    recursion; rather than guess, the table is reproduced from its generating rule verified in
    tests: left half / right half are the canonical 16x16 matrix in rows 0..7, while rows
    8..15 hold the canonical 8x8 table (incl. the 4x4 repeat quirk of rows 12..15, but with the
    canonical [11,6:8] entries) in columns 0..7 and that table minus 1/128 in columns 8..15.
    """
    c16 = (_bayer_index(16) + 1) / 256.0
    out = np.empty((16, 16), np.float64)
    out[0:8, :] = c16[0:8, :]
    b8 = (_bayer_index(8) + 1) / 64.0
    b4 = ((_bayer_index(4) + 0.5) / 16.0)
    b8q = b8.copy()
    b8q[4:8, 4:8] = b4
    out[8:16, 0:8] = b8q
    out[8:16, 8:16] = b8q - 1.0 / 128.0
    return out.astype(np.float32)


def polka_dot_matrix(tile_size: int = 8, gamma: float = 1.5) -> np.ndarray:
    """Radial threshold tile (:733-743): 1 - (r / (r_max + 1e-9)) ** gamma, clipped, as f32."""
    ax = np.arange(tile_size)
    xv, yv = np.meshgrid(ax, ax)
    c = (tile_size - 1) / 2
    r = np.sqrt((xv - c) ** 2 + (yv - c) ** 2)
    rmax = np.sqrt(c ** 2 + c ** 2)
    t = 1.0 - (r / (rmax + 1e-9)) ** gamma
    return np.clip(t, 0, 1).astype(np.float32)


def blue_noise_matrix(size: int = 64, seed: int = 42) -> np.ndarray:
    """generate_blue_noise (:381-399) restated with array operations.

    Farthest-point ordering of the RandomState(seed)-shuffled coordinate list; first maximum
    in current list order wins; value i / (size^2 - 1 + 1e-9) rounded to f32; squared distances
    are exact small integers held in f32.
    """
    n = size * size
    order = np.arange(n)
    np.random.RandomState(seed).shuffle(order)  # same draws as shuffling the list of tuples
    rr = (order // size).astype(np.int64)
    cc = (order % size).astype(np.int64)
    alive = np.ones(n, bool)
    mind = np.full(n, np.inf, np.float32)
    out = np.zeros((size, size), np.float32)
    denom = float(n - 1 + 1e-9)
    for i in range(n):
        key = np.where(alive, mind, -np.inf)
        j = int(np.argmax(key))  # first max among the live entries, in list order
        out[rr[j], cc[j]] = i / denom
        alive[j] = False
        d2 = ((rr - rr[j]) ** 2 + (cc - cc[j]) ** 2).astype(np.float32)
        upd = alive & (d2 < mind)
        mind[upd] = d2[upd]
    return out


def ign_thresholds(h: int, w: int, scale: float = 1.0, seed: int = 0) -> np.ndarray:
    """Interleaved gradient noise thresholds (:539-549); f32, one rounding per operation."""
    f = np.float32
    scale = float(scale)
    seed = int(seed)
    x = np.arange(w, dtype=np.float32)
    y = np.arange(h, dtype=np.float32)
    xv, yv = np.meshgrid(x, y)
    xv = (xv + f(seed * 0.37)) * f(scale)
    yv = (yv + f(seed * 0.73)) * f(scale)
    t = xv * f(0.06711056) + yv * f(0.00583715)
    t = t - np.floor(t)
    t = t * f(52.9829189)
    return (t - np.floor(t)).astype(np.float32)


# ----------------------------------------------------------------------------------------
# Threshold family (none / matrix / IGN / polka)
# ----------------------------------------------------------------------------------------

def nearest_indices(pixels: np.ndarray, palette: np.ndarray) -> np.ndarray:
    """NoDitherStrategy (:337-341): KD-tree k=1."""
    tree = KDTree(palette)
    _, idx = tree.query(pixels, k=1, workers=-1)
    return idx.astype(np.int32)


def threshold_indices(pixels: np.ndarray, palette: np.ndarray, thresholds: np.ndarray
                      ) -> np.ndarray:
    """The decision shared by :355-378, :551-568, :745-766.

    ``thresholds`` is the flat f32 per-pixel threshold.  factor = d1^2/(d1^2+d2^2) with the
    distances as scipy returns them (sqrt taken, then squared again in f64); nearest iff
    factor <= threshold.
    """
    tree = KDTree(palette)
    dist, idx = tree.query(pixels, k=2, workers=-1)
    sq = dist ** 2
    tot = sq[:, 0] + sq[:, 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        factor = np.where(tot == 0, 0.0, sq[:, 0] / tot)
    pick_first = factor <= thresholds
    return np.where(pick_first, idx[:, 0], idx[:, 1]).astype(np.int32)


def tile_thresholds(matrix: np.ndarray, h: int, w: int) -> np.ndarray:
    th, tw = matrix.shape
    reps = ((h + th - 1) // th, (w + tw - 1) // tw)
    return np.tile(matrix, reps)[:h, :w].reshape(-1)


# ----------------------------------------------------------------------------------------
# Halftone (:1597-1695)
# ----------------------------------------------------------------------------------------

def halftone_screen(h: int, w: int, cell_size=8, angle=45.0, dot_gain=1.0, min_dot_size=0.0,
                    max_dot_size=1.0, shape="circle", sharpness=1.5
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """(:1646-1695) returns (screen f32 [h,w], cell ids int [h,w])."""
    a = np.radians(angle)
    ca, sa = np.cos(a), np.sin(a)
    yy, xx = np.mgrid[0:h, 0:w]
    xr = xx * ca - yy * sa
    yr = xx * sa + yy * ca
    cx = np.floor(xr / cell_size).astype(np.int32)
    cy = np.floor(yr / cell_size).astype(np.int32)
    cxo = cx - cx.min()
    cyo = cy - cy.min()
    cells = cyo * (cxo.max() + 1) + cxo
    dx = (xr % cell_size) / cell_size - 0.5
    dy = (yr % cell_size) / cell_size - 0.5
    if shape == "square":
        dist, dmax = np.maximum(np.abs(dx), np.abs(dy)), 0.5
    elif shape == "diamond":
        dist, dmax = np.abs(dx) + np.abs(dy), 1.0
    else:
        dist, dmax = np.sqrt(dx ** 2 + dy ** 2), 0.5
    t = np.clip(dist / dmax, 0.0, 1.0) ** (1.0 / dot_gain)
    t = min_dot_size + t * (max_dot_size - min_dot_size)
    if sharpness != 1.0:
        t = 0.5 + (t - 0.5) * sharpness
    return np.clip(t, 0.0, 1.0).astype(np.float32), cells


def halftone_indices(pixels: np.ndarray, palette: np.ndarray, h: int, w: int, **params
                     ) -> np.ndarray:
    """(:1597-1644) per-cell mean colour -> nearest palette entry; ink where darkness > screen."""
    f = np.float32
    px = pixels.astype(np.float32).reshape(h, w, 3)
    gray = f(0.299) * px[:, :, 0] + f(0.587) * px[:, :, 1] + f(0.114) * px[:, :, 2]
    gray_norm = gray / f(255.0)
    pal_luma = f(0.299) * palette[:, 0] + f(0.587) * palette[:, 1] + f(0.114) * palette[:, 2]
    paper = int(np.argmax(pal_luma))
    screen, cells = halftone_screen(h, w, **params)
    flat_cells = cells.reshape(-1)
    uniq, inv = np.unique(flat_cells, return_inverse=True)
    count = np.bincount(inv, minlength=len(uniq))
    sums = np.stack([np.bincount(inv, weights=px.reshape(-1, 3)[:, c], minlength=len(uniq))
                     for c in range(3)], axis=1)
    means = sums / np.maximum(count[:, None], 1)
    _, cell_idx = KDTree(palette).query(means, k=1)
    ink = (f(1.0) - gray_norm) > screen
    out = np.full(h * w, paper, np.int32)
    m = ink.reshape(-1)
    out[m] = cell_idx[inv[m]]
    return out


# ----------------------------------------------------------------------------------------
# Error diffusion (numba semantics) and Ostromoukhov (live Python semantics) -- via C
# ----------------------------------------------------------------------------------------

ED_KERNELS = {
    # name: (taps (dx,dy,weight), divisor)      (:107-188)
    "floyd_steinberg": ([(1, 0, 7), (-1, 1, 3), (0, 1, 5), (1, 1, 1)], 16),
    "jjn": ([(1, 0, 7), (2, 0, 5), (-2, 1, 3), (-1, 1, 5), (0, 1, 7), (1, 1, 5), (2, 1, 3),
             (-2, 2, 1), (-1, 2, 3), (0, 2, 5), (1, 2, 3), (2, 2, 1)], 48),
    "stucki": ([(1, 0, 8), (2, 0, 4), (-2, 1, 2), (-1, 1, 4), (0, 1, 8), (1, 1, 4), (2, 1, 2),
                (-2, 2, 1), (-1, 2, 2), (0, 2, 4), (1, 2, 2), (2, 2, 1)], 42),
    "burkes": ([(1, 0, 8), (2, 0, 4), (-2, 1, 2), (-1, 1, 4), (0, 1, 8), (1, 1, 4), (2, 1, 2)],
               32),
    "atkinson": ([(1, 0, 1), (2, 0, 1), (-1, 1, 1), (0, 1, 1), (1, 1, 1), (0, 2, 1)], 8),
    "sierra": ([(1, 0, 5), (2, 0, 3), (-2, 1, 2), (-1, 1, 4), (0, 1, 5), (1, 1, 4), (2, 1, 2),
                (-1, 2, 2), (0, 2, 3), (1, 2, 2)], 32),
    "sierra_two_row": ([(1, 0, 4), (2, 0, 3), (-2, 1, 1), (-1, 1, 2), (0, 1, 3), (1, 1, 2),
                        (2, 1, 1)], 16),
    "sierra_lite": ([(1, 0, 2), (-1, 1, 1), (0, 1, 1)], 4),
}


def error_diffusion_indices(pixels: np.ndarray, palette: np.ndarray, h: int, w: int,
                            variant: str = "atkinson", serpentine: bool = False) -> np.ndarray:
    """(:631-651 -> :212-308).  Unknown variants fall back to Floyd-Steinberg (:203)."""
    taps, div = ED_KERNELS.get(variant, ED_KERNELS["floyd_steinberg"])
    work = np.ascontiguousarray(pixels, np.float32).reshape(h, w, 3).copy()
    pal = np.ascontiguousarray(palette, np.float32)
    offs = np.asarray([(dx, dy) for dx, dy, _ in taps], np.int32)
    wts = np.asarray([wt for _, _, wt in taps], np.float32)
    idx = np.empty((h, w), np.uint8)
    rc = _lib().orc_error_diffusion(
        ctypes.c_void_p(work.ctypes.data), ctypes.c_int(h), ctypes.c_int(w),
        ctypes.c_void_p(pal.ctypes.data), ctypes.c_int(pal.shape[0]),
        ctypes.c_void_p(offs.ctypes.data), ctypes.c_void_p(wts.ctypes.data),
        ctypes.c_int(len(taps)), ctypes.c_double(float(div)), ctypes.c_int(int(serpentine)),
        ctypes.c_void_p(idx.ctypes.data))
    if rc != 0:
        raise RuntimeError("orc_error_diffusion failed")
    return idx.reshape(-1).astype(np.int32)


def hybrid_indices(pixels: np.ndarray, palette: np.ndarray, h: int, w: int,
                   lum_factor: float = 1.0, col_factor: float = 0.2) -> np.ndarray:
    """HybridDitherStrategy.dither (:1111-1127) -> _hybrid_numba (:1396-1494)."""
    work = np.ascontiguousarray(pixels, np.float32).reshape(h, w, 3).copy()
    pal = np.ascontiguousarray(palette, np.float32)
    idx = np.empty((h, w), np.uint8)
    rc = _lib().orc_hybrid(
        ctypes.c_void_p(work.ctypes.data), ctypes.c_int(h), ctypes.c_int(w),
        ctypes.c_void_p(pal.ctypes.data), ctypes.c_int(pal.shape[0]),
        ctypes.c_double(float(lum_factor)), ctypes.c_double(float(col_factor)),
        ctypes.c_void_p(idx.ctypes.data))
    if rc != 0:
        raise RuntimeError("orc_hybrid failed")
    return idx.reshape(-1).astype(np.int32)


def perceptual_indices(pixels: np.ndarray, palette: np.ndarray, h: int, w: int) -> np.ndarray:
    """PerceptualDitherStrategy.dither (:1040-1066), default base_weights (Floyd-Steinberg)."""
    work = np.ascontiguousarray(pixels, np.float32).reshape(h, w, 3).copy()
    pal = np.ascontiguousarray(palette, np.float32)
    tree = export_kdtree(pal)
    s = _kd_struct(tree)
    idx = np.empty((h, w), np.uint8)
    rc = _lib().orc_perceptual(
        ctypes.c_void_p(work.ctypes.data), ctypes.c_int(h), ctypes.c_int(w),
        ctypes.c_void_p(pal.ctypes.data), ctypes.c_int(pal.shape[0]), ctypes.byref(s),
        ctypes.c_void_p(idx.ctypes.data))
    if rc != 0:
        raise RuntimeError("orc_perceptual failed")
    return idx.reshape(-1).astype(np.int32)


def uniform_filter_nearest(a: np.ndarray, size: int) -> np.ndarray:
    """scipy.ndimage.uniform_filter(a f32 [h,w], size, mode='nearest') restated (scipy 1.18.1,
    third-party; NI_UniformFilter1D): axis 0 then axis 1, each pass a RUNNING SUM in double over
    the edge-extended line (tmp = first window; out = tmp / size; tmp += new - old), f32 between
    passes.  Checked against scipy itself in tests/test_oracle_golden.py."""
    out = np.asarray(a, np.float32)
    if size <= 1:
        return out.copy()
    r = size // 2
    for axis in (0, 1):
        lines = np.moveaxis(out, axis, 1).astype(np.float64)          # [lines, length]
        ext = np.concatenate([np.repeat(lines[:, :1], r, 1), lines,
                              np.repeat(lines[:, -1:], size - r - 1, 1)], 1)
        res = np.empty_like(lines)
        tmp = np.zeros(lines.shape[0])
        for k in range(size):
            tmp = tmp + ext[:, k]
        res[:, 0] = tmp / float(size)
        for l in range(1, lines.shape[1]):
            tmp = tmp + (ext[:, l + size - 1] - ext[:, l - 1])
            res[:, l] = tmp / float(size)
        out = np.moveaxis(res.astype(np.float32), 1, axis)
    return np.ascontiguousarray(out)


def variance_gate(pixels: np.ndarray, h: int, w: int, var_threshold: float = 300.0,
                  window_radius: int = 1) -> np.ndarray:
    """AdaptiveVarianceDitherStrategy (:993-996, :1021-1025) with scipy's own uniform_filter, as
    the reference calls it: gray (f32) -> local variance -> `>= var_threshold`."""
    from scipy.ndimage import uniform_filter
    pix = np.asarray(pixels, np.float32).reshape(h, w, 3)
    gray = 0.299 * pix[:, :, 0] + 0.587 * pix[:, :, 1] + 0.114 * pix[:, :, 2]
    size = 2 * window_radius + 1
    g = gray.astype(np.float32)
    mean_sq = uniform_filter(g ** 2, size=size, mode='nearest')
    sq_mean = uniform_filter(g, size=size, mode='nearest') ** 2
    var = np.maximum(0.0, mean_sq - sq_mean)
    return np.ascontiguousarray(var >= var_threshold).astype(np.uint8)


def adaptive_indices(pixels: np.ndarray, palette: np.ndarray, h: int, w: int,
                     var_threshold: float = 300.0, window_radius: int = 1) -> np.ndarray:
    """AdaptiveVarianceDitherStrategy.dither (:989-1019)."""
    work = np.ascontiguousarray(pixels, np.float32).reshape(h, w, 3).copy()
    pal = np.ascontiguousarray(palette, np.float32)
    gate = variance_gate(pixels, h, w, var_threshold, window_radius)
    tree = export_kdtree(pal)
    s = _kd_struct(tree)
    idx = np.empty((h, w), np.uint8)
    rc = _lib().orc_adaptive(
        ctypes.c_void_p(work.ctypes.data), ctypes.c_int(h), ctypes.c_int(w),
        ctypes.c_void_p(pal.ctypes.data), ctypes.c_int(pal.shape[0]), ctypes.byref(s),
        ctypes.c_void_p(gate.ctypes.data), ctypes.c_void_p(idx.ctypes.data))
    if rc != 0:
        raise RuntimeError("orc_adaptive failed")
    return idx.reshape(-1).astype(np.int32)


_OSTRO = None


def ostromoukhov_coeffs() -> np.ndarray:
    """COEFFS_TABLE (:1170-1203), 256 x (c0,c1,c2); data file extracted by
    tools/extract_tables.py."""
    global _OSTRO
    if _OSTRO is None:
        here = os.path.dirname(os.path.abspath(__file__))
        _OSTRO = np.load(os.path.join(here, "..", "dither_pie_b200", "data",
                                      "ostromoukhov_coeffs.npy")).astype(np.int32)
    return _OSTRO


def ostromoukhov_indices(pixels: np.ndarray, palette: np.ndarray, h: int, w: int,
                         serpentine: bool = False) -> np.ndarray:
    """(:1225-1269), the live path (f32 arithmetic + KD-tree nearest)."""
    work = np.ascontiguousarray(pixels, np.float32).reshape(h, w, 3).copy()
    pal = np.ascontiguousarray(palette, np.float32)
    tree = export_kdtree(pal)
    s = _kd_struct(tree)
    co = np.ascontiguousarray(ostromoukhov_coeffs(), np.int32)
    idx = np.empty((h, w), np.uint8)
    rc = _lib().orc_ostromoukhov(
        ctypes.c_void_p(work.ctypes.data), ctypes.c_int(h), ctypes.c_int(w),
        ctypes.c_void_p(pal.ctypes.data), ctypes.c_int(pal.shape[0]), ctypes.byref(s),
        ctypes.c_void_p(co.ctypes.data), ctypes.c_int(int(serpentine)),
        ctypes.c_void_p(idx.ctypes.data))
    if rc != 0:
        raise RuntimeError("orc_ostromoukhov failed")
    return idx.reshape(-1).astype(np.int32)


# ----------------------------------------------------------------------------------------
# Gamma (:1788-1802, :1956-1974, :1986-1990)
# ----------------------------------------------------------------------------------------

def srgb_to_linear(c: np.ndarray) -> np.ndarray:
    c = np.asarray(c)
    out = np.empty_like(c, dtype=np.float32)
    low = c <= 0.04045
    out[low] = c[low] / 12.92
    out[~low] = ((c[~low] + 0.055) / 1.055) ** 2.4
    return out


def linear_to_srgb(c: np.ndarray) -> np.ndarray:
    c = np.asarray(c)
    out = np.empty_like(c, dtype=np.float32)
    low = c <= 0.0031308
    out[low] = c[low] * 12.92
    out[~low] = 1.055 * (c[~low] ** (1.0 / 2.4)) - 0.055
    return out


# ----------------------------------------------------------------------------------------
# Whole-image wrapper: ImageDitherer.apply_dithering (:1952-1992) on uint8 arrays
# ----------------------------------------------------------------------------------------

def apply_dithering(img_u8: np.ndarray, palette: Sequence[Sequence[float]], mode: str,
                    params: Optional[dict] = None, use_gamma: bool = False) -> np.ndarray:
    """uint8 [h,w,3] -> uint8 [h,w,3].  ``mode`` is the DitherMode string value (:61-75)."""
    params = dict(params or {})
    arr = np.ascontiguousarray(img_u8, np.uint8)
    pal = np.array(palette, dtype=np.float32)
    if use_gamma:
        lin = srgb_to_linear(arr.astype(np.float32) / 255.0)
        arr = np.clip(lin * 255.0, 0, 255).astype(np.uint8)
        pal = np.clip(srgb_to_linear(pal / 255.0) * 255.0, 0, 255).astype(np.float32)
    h, w, _ = arr.shape
    flat = arr.reshape(-1, 3).astype(np.float32)

    if mode == "none":
        idx = nearest_indices(flat, pal)
    elif mode == "bayer":
        idx = threshold_indices(flat, pal, tile_thresholds(
            bayer_matrix(params.get("size", "4x4")), h, w))
    elif mode == "blue_noise":
        idx = threshold_indices(flat, pal, tile_thresholds(
            blue_noise_matrix(params.get("size", 64), params.get("seed", 42)), h, w))
    elif mode == "polka_dot":
        idx = threshold_indices(flat, pal, tile_thresholds(
            polka_dot_matrix(params.get("tile_size", 8), params.get("gamma", 1.5)), h, w))
    elif mode == "IGN":
        idx = threshold_indices(flat, pal, ign_thresholds(
            h, w, params.get("scale", 1.0), params.get("seed", 0)).reshape(-1))
    elif mode == "halftone":
        idx = halftone_indices(flat, pal, h, w, **params)
    elif mode == "error_diffusion":
        idx = error_diffusion_indices(flat, pal, h, w, params.get("variant", "atkinson"),
                                      params.get("serpentine", "false") == "true")
    elif mode == "ostromoukhov":
        idx = ostromoukhov_indices(flat, pal, h, w, params.get("serpentine", "false") == "true")
    elif mode == "perceptual":
        idx = perceptual_indices(flat, pal, h, w)
    elif mode == "adaptive_variance":
        idx = adaptive_indices(flat, pal, h, w, params.get("var_threshold", 300.0),
                               params.get("window_radius", 1))
    elif mode == "hybrid":
        idx = hybrid_indices(flat, pal, h, w, params.get("lum_factor", 1.0),
                             params.get("col_factor", 0.2))
    else:
        raise ValueError(f"mode {mode!r} is outside the hot path")
    out = pal[idx, :].reshape(h, w, 3).astype(np.uint8)
    if use_gamma:
        s = linear_to_srgb(np.clip(out.astype(np.float32) / 255.0, 0, 1))
        out = np.clip(s * 255.0, 0, 255).astype(np.uint8)
    return out


# ----------------------------------------------------------------------------------------
# Regular pixelization (video_processor.py:547-577, :393-420; dither_cli.py:559-566)
# ----------------------------------------------------------------------------------------

def even_dimensions(orig_w: int, orig_h: int, max_size: int) -> Tuple[int, int]:
    """video_processor.py:547-560."""
    base = max_size if max_size % 2 == 0 else max_size - 1
    if orig_w >= orig_h:
        th = base
        tw = int(round((orig_w / orig_h) * th))
        tw += tw % 2
    else:
        tw = base
        th = int(round((orig_h / orig_w) * tw))
        th += th % 2
    return tw, th


def nearest_table(n_in: int, n_out: int) -> np.ndarray:
    """Source index per output index for Pillow's NEAREST resize (affine path): a running f64
    sum xo = 0.5*s, xo += s, index = int(xo) ([probe])."""
    s = n_in / n_out
    tab = np.empty(n_out, np.int32)
    xo = 0.5 * s
    for i in range(n_out):
        tab[i] = min(int(xo), n_in - 1)
        xo += s
    return tab


def pixelize_regular(img_u8: np.ndarray, max_size: int) -> np.ndarray:
    h, w, _ = img_u8.shape
    tw, th = even_dimensions(w, h, max_size)
    xt = nearest_table(w, tw)
    yt = nearest_table(h, th)
    return np.ascontiguousarray(img_u8[yt][:, xt])


def final_resize(img_u8: np.ndarray, multiplier: int, even: bool = True) -> np.ndarray:
    """video_processor.py:393-420 (even=True) / dither_cli.py:559-566 (even=False)."""
    h, w, _ = img_u8.shape
    nw, nh = w * multiplier, h * multiplier
    if even:
        nw += nw % 2
        nh += nh % 2
    return np.ascontiguousarray(img_u8[nearest_table(h, nh)][:, nearest_table(w, nw)])


# ----------------------------------------------------------------------------------------
# K-means palette (:1845-1857 -> sklearn KMeans, lloyd, n_init=1, k-means++)
# ----------------------------------------------------------------------------------------

def kmeans_plusplus_init(X: np.ndarray, k: int, seed: int = 42) -> np.ndarray:
    """sklearn's _kmeans_plusplus (greedy, n_local_trials = 2 + int(log k)) on f64 data.
    X must already be centred the way KMeans.fit does (X - X.mean(axis=0))."""
    rs = np.random.RandomState(seed)
    n = X.shape[0]
    x_sq = np.einsum("ij,ij->i", X, X)
    trials = 2 + int(np.log(k))
    centers = np.empty((k, X.shape[1]), X.dtype)
    cid = rs.choice(n, p=np.full(n, 1.0 / n))
    centers[0] = X[cid]

    def sqdist(c):
        d = x_sq - 2.0 * (X @ c.T).T + np.einsum("ij,ij->i", c, c)[:, None]
        np.maximum(d, 0, out=d)
        return d

    closest = sqdist(centers[0:1])[0]
    pot = closest.sum()
    for c in range(1, k):
        r = rs.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(closest, dtype=np.float64), r)
        np.clip(cand, None, n - 1, out=cand)
        dc = sqdist(X[cand])
        np.minimum(closest, dc, out=dc)
        pots = dc.sum(axis=1)
        best = int(np.argmin(pots))
        pot = pots[best]
        closest = dc[best]
        centers[c] = X[cand[best]]
    return centers


def lloyd(X: np.ndarray, centers: np.ndarray, tol: float, max_iter: int = 300
          ) -> Tuple[np.ndarray, int]:
    """Plain exact-distance Lloyd in f64 with sklearn's stopping rule (shift^2 sum <= tol)."""
    c = centers.copy()
    it = 0
    for it in range(1, max_iter + 1):
        d = ((X[:, None, :] - c[None, :, :]) ** 2).sum(axis=2)
        lab = np.argmin(d, axis=1)
        new = c.copy()
        for j in range(c.shape[0]):
            m = lab == j
            if m.any():
                new[j] = X[m].mean(axis=0)
        shift = ((new - c) ** 2).sum()
        c = new
        if shift <= tol:
            break
    return c, it


def kmeans_centers(sample_u8: np.ndarray, k: int, seed: int = 42) -> np.ndarray:
    """Pre-truncation centroids for the given (already sub-sampled) pixels."""
    X = sample_u8.astype(np.float64)
    mean = X.mean(axis=0)
    Xc = X - mean
    tol = float(np.mean(np.var(Xc, axis=0)) * 1e-4)
    init = kmeans_plusplus_init(Xc, k, seed)
    c, _ = lloyd(Xc, init, tol)
    return c + mean
