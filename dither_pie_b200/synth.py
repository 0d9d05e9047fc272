"""Deterministic synthetic frames and palettes.

Shared by tests/, bench.py and tools/make_golden.py so that every side sees the same bytes.
"""
from __future__ import annotations

import numpy as np


def frame(h: int, w: int, seed: int = 0) -> np.ndarray:
    """Gradient + iid integer noise uniform on [-16, 16]; uint8 [h, w, 3]."""
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    r = np.broadcast_to(x * 255 // max(w - 1, 1), (h, w))
    g = np.broadcast_to(y * 255 // max(h - 1, 1), (h, w))
    b = (x + y) * 255 // max(w + h - 2, 1)
    base = np.stack([r, g, b], axis=2)
    noise = np.random.RandomState(seed).randint(-16, 17, size=(h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def noise_frame(h: int, w: int, seed: int = 0) -> np.ndarray:
    """iid uniform uint8."""
    return np.random.RandomState(seed).randint(0, 256, size=(h, w, 3)).astype(np.uint8)


def blocks_frame(h: int, w: int, seed: int = 0, block: int = 16, levels: int = 0) -> np.ndarray:
    """Constant tiles of random colours (tie-heavy).  levels>0 snaps colours to a coarse
    lattice (multiples of 255/(levels-1)) which makes exact distance ties common."""
    rs = np.random.RandomState(seed)
    bh, bw = (h + block - 1) // block, (w + block - 1) // block
    if levels > 1:
        step = 255 // (levels - 1)
        cols = rs.randint(0, levels, size=(bh, bw, 3)) * step
    else:
        cols = rs.randint(0, 256, size=(bh, bw, 3))
    img = np.repeat(np.repeat(cols, block, axis=0), block, axis=1)[:h, :w]
    return img.astype(np.uint8)


def random_palette(k: int, seed: int = 2024) -> np.ndarray:
    """First k unique rows of RandomState(seed).randint(0,256,(.,3)); int64 [k,3]."""
    rs = np.random.RandomState(seed)
    rows = rs.randint(0, 256, size=(4 * k + 64, 3))
    seen, out = set(), []
    for r in rows:
        t = (int(r[0]), int(r[1]), int(r[2]))
        if t not in seen:
            seen.add(t)
            out.append(t)
        if len(out) == k:
            break
    return np.asarray(out, dtype=np.int64)


def lattice_palette(k: int, seed: int = 7, step: int = 51) -> np.ndarray:
    """k unique colours on a coarse lattice (multiples of ``step``) -- tie-heavy."""
    rs = np.random.RandomState(seed)
    n = 255 // step + 1
    seen, out = set(), []
    while len(out) < k:
        t = tuple(int(v) * step for v in rs.randint(0, n, size=3))
        if t not in seen:
            seen.add(t)
            out.append(t)
    return np.asarray(out, dtype=np.int64)


# a few entries of the reference's palette.json (data, utils.py:31-50 format), used by tests
PICO8 = ["#000000", "#5f574f", "#c2c3c7", "#fff1e8", "#ff004d", "#ffa300", "#ffec27", "#00e436", "#29adff", "#83769c", "#ff77a8", "#ffccaa", "#1d2b53", "#7e253b", "#008751", "#ab5236"]
C64 = ["#000000", "#ffffff", "#880000", "#aaffee", "#cc44cc", "#00cc55", "#0000aa", "#e6e600", "#dd8855", "#664400", "#ff7777", "#333333", "#777777", "#aaff66", "#00aaff", "#bbbbbb"]
GB_POCKET = ["#000000", "#555555", "#aaaaaa", "#ffffff"]


def hex_palette(colors) -> np.ndarray:
    return np.asarray([[int(c[i:i + 2], 16) for i in (1, 3, 5)] for c in colors], dtype=np.int64)
