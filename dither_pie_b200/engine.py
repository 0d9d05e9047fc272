"""Host-side engine over the C ABI: palette handles, threshold sources, geometry tables and the
per-mode dispatch on DEVICE pointers.  Everything per-pixel happens in libditherpie_b200.so;
what is computed here is set-up that the reference also computes once per strategy object
(threshold matrices, :402-448, :451-499, :733-743; gamma LUTs :1788-1802; the KD-tree :358).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import DeviceBuffer, Geometry, check, lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# ----------------------------------------------------------------------------------------
# threshold sources (host set-up, cached like the reference's class-level caches)
# ----------------------------------------------------------------------------------------


def _bayer_index(n: int) -> np.ndarray:
    m = np.zeros((1, 1), np.int64)
    while m.shape[0] < n:
        m = np.block([[4 * m, 4 * m + 2], [4 * m + 3, 4 * m + 1]])
    return m


def bayer_matrix(size: str) -> np.ndarray:
    """DitherUtils.BAYER* / PSX4x4 (dithering_lib.py:1705-1768), generated instead of typed in.
    The reference's departures from the textbook recursion are reproduced deliberately: 8x8 has
    its lower-right 4x4 block replaced by the 4x4 table and [3,6:8] = 0.84375, 0.34375; 16x16
    is the textbook matrix in rows 0-7 and two copies of an 8x8 variant in rows 8-15."""
    b4 = (_bayer_index(4) + 0.5) / 16.0
    if size == "2x2":
        m = (_bayer_index(2) + 1) / 4.0
    elif size in ("psx4x4", "psx"):
        m = np.array([[1, 9, 3, 11], [13, 5, 15, 7], [3, 11, 1, 9], [15, 7, 13, 5]]) / 16.0
    elif size == "8x8":
        m = (_bayer_index(8) + 1) / 64.0
        m[4:8, 4:8] = b4
        m[3, 6], m[3, 7] = 0.84375, 0.34375
    elif size == "16x16":
        m = np.empty((16, 16))
        m[0:8, :] = ((_bayer_index(16) + 1) / 256.0)[0:8, :]
        q = (_bayer_index(8) + 1) / 64.0
        q[4:8, 4:8] = b4
        m[8:16, 0:8] = q
        m[8:16, 8:16] = q - 1.0 / 128.0
    else:
        m = b4  # unknown sizes fall back to 4x4 (:441-442)
    return np.ascontiguousarray(m, dtype=np.float32)


def polka_dot_matrix(tile_size: int, gamma: float) -> np.ndarray:
    """PolkaDotDitherStrategy._generate_polka_dot_matrix (:733-743)."""
    ax = np.arange(tile_size)
    xv, yv = np.meshgrid(ax, ax)
    c = (tile_size - 1) / 2
    norm = np.sqrt((xv - c) ** 2 + (yv - c) ** 2) / (np.sqrt(c ** 2 + c ** 2) + 1e-9)
    return np.clip(1.0 - norm ** gamma, 0, 1).astype(np.float32)


_blue_cache: Dict[Tuple[int, int], np.ndarray] = {}
_blue_lock = threading.Lock()


def blue_noise_matrix(size: int = 64, seed: int = 42) -> np.ndarray:
    """generate_blue_noise (:381-399): farthest-point ordering of a shuffled coordinate list
    (numpy's RandomState shuffle, as the reference), replayed natively in integer arithmetic
    (dp_blue_noise_from_order, host code in the library; the reference is a pure-Python O(n^2)
    double loop: 7 s at size 64)."""
    key = (int(size), int(seed))
    with _blue_lock:
        if key in _blue_cache:
            return _blue_cache[key]
    n = size * size
    order = np.arange(n)
    np.random.RandomState(seed).shuffle(order)
    order = np.ascontiguousarray(order, np.int32)
    out = np.zeros((size, size), np.float32)
    check(_capi.lib().dp_blue_noise_from_order(order.ctypes.data, int(size), out.ctypes.data),
          "dp_blue_noise_from_order")
    with _blue_lock:
        _blue_cache[key] = out
    return out


def ostromoukhov_coeffs() -> np.ndarray:
    """COEFFS_TABLE (:1170-1203) as int32 [256,3]."""
    return np.load(os.path.join(_DATA, "ostromoukhov_coeffs.npy")).astype(np.int32)


# ----------------------------------------------------------------------------------------
# gamma (:1788-1802)
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


def gamma_in_lut() -> np.ndarray:
    """uint8 sRGB -> uint8 linear, exactly the element-wise chain of :1957-1959."""
    v = np.arange(256, dtype=np.uint8).astype(np.float32) / 255.0
    return np.clip(srgb_to_linear(v) * 255.0, 0, 255).astype(np.uint8)


def gamma_out_lut() -> np.ndarray:
    """uint8 linear -> uint8 sRGB, the chain of :1987-1989."""
    v = np.arange(256, dtype=np.uint8).astype(np.float32) / 255.0
    return np.clip(linear_to_srgb(np.clip(v, 0, 1)) * 255.0, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------
# palette handle
# ----------------------------------------------------------------------------------------

def export_kdtree(palette_f32: np.ndarray):
    """scipy.spatial.KDTree(palette) flattened in pre-order.  The tree is built
    by the installed scipy so the in-leaf order -- which decides exact ties -- is the
    reference's by construction."""
    from scipy.spatial import KDTree, cKDTree

    tree = KDTree(palette_f32)
    root = cKDTree.tree.__get__(tree)
    sd, sp, st, en, le, gr = [], [], [], [], [], []
    stack = [(root, -1, False)]
    while stack:
        node, parent, is_greater = stack.pop()
        me = len(sd)
        sd.append(int(node.split_dim))
        sp.append(float(node.split))
        st.append(int(node.start_idx))
        en.append(int(node.end_idx))
        le.append(-1)
        gr.append(-1)
        if parent >= 0:
            (gr if is_greater else le)[parent] = me
        if node.split_dim != -1:
            stack.append((node.greater, me, True))
            stack.append((node.lesser, me, False))
    return (np.asarray(sd, np.int32), np.asarray(sp, np.float64), np.asarray(st, np.int32),
            np.asarray(en, np.int32), np.asarray(le, np.int32), np.asarray(gr, np.int32),
            np.ascontiguousarray(tree.indices, np.int32),
            np.ascontiguousarray(tree.mins, np.float64),
            np.ascontiguousarray(tree.maxes, np.float64))


class PaletteHandle:
    """Device-resident palette (+ KD-tree, gamma LUT, candidate grid); owns a dp_palette."""

    def __init__(self, palette: Sequence[Sequence[float]], use_gamma: bool = False,
                 search_space: bool = False):
        """``palette``: rows as the caller's ImageDitherer holds them (sRGB 0..255).
        ``search_space=True`` means the rows are already in the space the search runs in (the
        strategy-level API hands over an already linearised ``palette_arr``)."""
        _capi.ensure_device()
        pal = np.array(palette, dtype=np.float32).reshape(-1, 3)
        self.use_gamma = bool(use_gamma)
        if use_gamma and not search_space:
            # (:1970-1974) palette -> linear, stays f32 (non-integral)
            pal = np.clip(srgb_to_linear(pal / 255.0) * 255.0, 0, 255).astype(np.float32)
        self.palette_f32 = np.ascontiguousarray(pal)
        # bytes written per row: palette_arr[idx].astype(uint8) (:1984) then linear->sRGB (:1986-1989)
        with np.errstate(invalid="ignore"):
            out = self.palette_f32.astype(np.uint8)
        in_lut = None
        if use_gamma:
            out = gamma_out_lut()[out]
            in_lut = gamma_in_lut()
        self.out_rgb = np.ascontiguousarray(out, np.uint8)
        self.K = int(pal.shape[0])
        kd = export_kdtree(self.palette_f32)
        h = C.c_void_p()
        check(lib().dp_palette_create(
            self.palette_f32.ctypes.data, self.K, self.out_rgb.ctypes.data,
            in_lut.ctypes.data if in_lut is not None else None,
            len(kd[0]), kd[0].ctypes.data, kd[1].ctypes.data, kd[2].ctypes.data,
            kd[3].ctypes.data, kd[4].ctypes.data, kd[5].ctypes.data, kd[6].ctypes.data,
            kd[7].ctypes.data, kd[8].ctypes.data, C.byref(h)), "dp_palette_create")
        self.handle = h.value

    def close(self):
        if getattr(self, "handle", None):
            lib().dp_palette_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_pal_cache: Dict[tuple, PaletteHandle] = {}
_pal_lock = threading.Lock()
_PAL_CACHE_MAX = 32


def get_palette(palette, use_gamma: bool = False, search_space: bool = False) -> PaletteHandle:
    """LRU-cached PaletteHandle (building one costs a scipy tree + a few device launches)."""
    arr = np.array(palette, dtype=np.float32).reshape(-1, 3)
    key = (arr.tobytes(), bool(use_gamma), bool(search_space), _capi.ensure_device())
    with _pal_lock:
        h = _pal_cache.pop(key, None)
        if h is not None:
            _pal_cache[key] = h
            return h
    h = PaletteHandle(arr, use_gamma, search_space)
    with _pal_lock:
        _pal_cache[key] = h
        while len(_pal_cache) > _PAL_CACHE_MAX:
            _pal_cache.pop(next(iter(_pal_cache)))
    return h


# ----------------------------------------------------------------------------------------
# device-side constant tables, cached per device
# ----------------------------------------------------------------------------------------

_tab_cache: Dict[tuple, DeviceBuffer] = {}
_tab_lock = threading.Lock()


_TAB_CACHE_MAX_BYTES = 64 << 20


def device_table(arr: np.ndarray) -> DeviceBuffer:
    """Small constant tables (threshold matrices, index tables) by content; LRU bounded by bytes."""
    arr = np.ascontiguousarray(arr)
    key = (arr.dtype.str, arr.shape, arr.tobytes(), _capi.ensure_device())
    with _tab_lock:
        buf = _tab_cache.pop(key, None)
        if buf is None:
            buf = DeviceBuffer(max(arr.nbytes, 4)).upload(arr)
            _capi.sync()
        _tab_cache[key] = buf
        total = sum(b.nbytes for b in _tab_cache.values())
        while total > _TAB_CACHE_MAX_BYTES and len(_tab_cache) > 1:
            k0 = next(iter(_tab_cache))
            total -= _tab_cache.pop(k0).nbytes
    return buf


# ----------------------------------------------------------------------------------------
# geometry: Pillow's NEAREST mapping (video_processor.py:547-577, :393-420)
# ----------------------------------------------------------------------------------------

def even_dimensions(orig_w: int, orig_h: int, max_size: int) -> Tuple[int, int]:
    """NeuralPixelizer._compute_even_dimensions (video_processor.py:547-560)."""
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
    """Pillow ImagingScaleAffine NEAREST: running double sum xo = 0.5*s; xo += s; int(xo)."""
    s = n_in / n_out
    tab = np.empty(n_out, np.int32)
    xo = 0.5 * s
    for i in range(n_out):
        tab[i] = min(int(xo), n_in - 1)
        xo += s
    return tab


# ----------------------------------------------------------------------------------------
# mode dispatch on device pointers
# ----------------------------------------------------------------------------------------

ED_VARIANTS = {"floyd_steinberg": 0, "jjn": 1, "stucki": 2, "burkes": 3, "atkinson": 4,
               "sierra": 5, "sierra_two_row": 6, "sierra_lite": 7}
HALFTONE_SHAPES = {"circle": 0, "square": 1, "diamond": 2}


def halftone_screen_host(h, w, cell_size, angle, dot_gain, min_dot_size, max_dot_size, shape,
                         sharpness) -> np.ndarray:
    """Host screen for dot_gain != 1 (needs pow(); :1646-1695).  Frame-invariant set-up."""
    a = np.radians(angle)
    ca, sa = np.cos(a), np.sin(a)
    yy, xx = np.mgrid[0:h, 0:w]
    xr = xx * ca - yy * sa
    yr = xx * sa + yy * ca
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
    return np.clip(t, 0.0, 1.0).astype(np.float32)


# Host-built halftone screens (dot_gain != 1 needs pow()): frame-invariant, so one device copy
# per (size, parameters, device), kept in an LRU bounded by total bytes (a 4K screen is 33 MB).
_screen_cache: "Dict[tuple, DeviceBuffer]" = {}
_screen_lock = threading.Lock()
_SCREEN_CACHE_MAX_BYTES = 512 << 20


def _halftone_screen_device(h: int, w: int, ht: dict) -> DeviceBuffer:
    key = (int(h), int(w), ht["cell_size"], ht["angle"], ht["dot_gain"], ht["min_dot_size"],
           ht["max_dot_size"], ht["shape"], ht["sharpness"], _capi.ensure_device())
    with _screen_lock:
        buf = _screen_cache.pop(key, None)
        if buf is not None:
            _screen_cache[key] = buf          # most recently used last
            return buf
    arr = halftone_screen_host(h, w, **ht)
    buf = DeviceBuffer(max(arr.nbytes, 4)).upload(arr)
    _capi.sync()
    with _screen_lock:
        _screen_cache[key] = buf
        total = sum(b.nbytes for b in _screen_cache.values())
        while total > _SCREEN_CACHE_MAX_BYTES and len(_screen_cache) > 1:
            k0 = next(iter(_screen_cache))
            total -= _screen_cache.pop(k0).nbytes   # dropped; freed when no Plan holds it any more
    return buf


class Plan:
    """Everything a (mode, params, geometry) needs on the device besides the palette.
    Built once, reused for every frame batch (the reference rebuilds its strategy per call,
    dithering_lib.py:1982; the result is the same because the strategy is stateless)."""

    def __init__(self, mode: str, params: Optional[dict], h: int, w: int,
                 src_hw: Optional[Tuple[int, int]] = None, upscale: int = 1):
        self.mode = mode
        self.params = dict(params or {})
        self.h, self.w = int(h), int(w)
        self.src_h, self.src_w = (src_hw if src_hw else (h, w))
        self.upscale = max(1, int(upscale or 1))
        self.out_h, self.out_w = self.h * self.upscale, self.w * self.upscale
        self.geo = Geometry()
        self.geo.src_h, self.geo.src_w = self.src_h, self.src_w
        self.geo.h, self.geo.w = self.h, self.w
        self.geo.upscale = self.upscale
        self._keep = []
        if (self.src_h, self.src_w) != (self.h, self.w):
            yt = device_table(nearest_table(self.src_h, self.h))
            xt = device_table(nearest_table(self.src_w, self.w))
            self._keep += [yt, xt]
            self.geo.ytab, self.geo.xtab = yt.ptr, xt.ptr
        p = self.params
        self.kind = None
        self.matrix = None
        self.mat_shape = (0, 0)
        self.ign = (0.0, 0.0, 1.0)
        if mode == "none":
            self.kind = 0
        elif mode in ("bayer", "blue_noise", "polka_dot"):
            if mode == "bayer":
                m = bayer_matrix(p.get("size", "4x4"))
            elif mode == "blue_noise":
                m = blue_noise_matrix(p.get("size", 64), p.get("seed", 42))
            else:
                m = polka_dot_matrix(p.get("tile_size", 8), p.get("gamma", 1.5))
            self.kind = 1
            self.matrix = device_table(m)
            self.mat_shape = m.shape
        elif mode == "IGN":
            scale, seed = float(p.get("scale", 1.0)), int(p.get("seed", 0))
            # python scalars are rounded to f32 before they meet the f32 arrays (:546-547)
            self.ign = (float(np.float32(seed * 0.37)), float(np.float32(seed * 0.73)),
                        float(np.float32(scale)))
            self.kind = 2
        elif mode == "halftone":
            self.ht = dict(cell_size=int(p.get("cell_size", 8)), angle=float(p.get("angle", 45.0)),
                           dot_gain=float(p.get("dot_gain", 1.0)),
                           min_dot_size=float(p.get("min_dot_size", 0.0)),
                           max_dot_size=float(p.get("max_dot_size", 1.0)),
                           shape=p.get("shape", "circle"), sharpness=float(p.get("sharpness", 1.5)))
            self.ht_screen = None
            if self.ht["dot_gain"] != 1.0:
                self.ht_screen = _halftone_screen_device(self.h, self.w, self.ht)
        elif mode == "error_diffusion":
            self.variant = ED_VARIANTS.get(p.get("variant", "atkinson"), 0)  # unknown -> FS (:203)
            self.serpentine = p.get("serpentine", "false") == "true"
        elif mode == "ostromoukhov":
            self.serpentine = p.get("serpentine", "false") == "true"
            self.coeffs = ostromoukhov_coeffs()
        elif mode == "adaptive_variance":
            # AdaptiveVarianceDitherStrategy.__init__ (:979-981)
            self.var_threshold = float(p.get("var_threshold", 300.0))
            self.window_radius = int(p.get("window_radius", 1))
        elif mode == "perceptual":
            pass   # PerceptualDitherStrategy (:1030-1066) with its default Floyd-Steinberg weights
        elif mode == "hybrid":
            # HybridDitherStrategy.__init__ (:1101-1104); dither() passes float(...) of both (:1121)
            self.lum_factor = float(p.get("lum_factor", 1.0))
            self.col_factor = float(p.get("col_factor", 0.2))
        else:
            raise ValueError(f"Unrecognized or out-of-scope dither mode: {mode!r}")
        self.fused_geometry = self.kind is not None

    # ---- run on device pointers -------------------------------------------------------
    def run(self, pal: PaletteHandle, src_ptr: int, frames: int, dst_ptr: int,
            idx_ptr: Optional[int] = None, stream=None):
        """src: u8 [frames, src_h, src_w, 3]; dst: u8 [frames, out_h, out_w, 3] (threshold
        family: pixelize/upscale fused; other modes need identity geometry).  ``dst_ptr`` may be
        None when ``idx_ptr`` (u8 [frames, h, w]) is given: index-plane-only output."""
        L = lib()
        if self.kind is not None:
            mat = self.matrix.ptr if self.matrix is not None else None
            check(L.dp_threshold_dither(pal.handle, src_ptr, frames, C.byref(self.geo), self.kind,
                                        mat, self.mat_shape[0], self.mat_shape[1],
                                        self.ign[0], self.ign[1], self.ign[2], dst_ptr, idx_ptr,
                                        stream), "dp_threshold_dither")
            return
        if (self.src_h, self.src_w) != (self.h, self.w) or self.upscale != 1:
            raise ValueError("this mode runs at identity geometry; resample separately")
        if self.mode == "halftone":
            a = np.radians(self.ht["angle"])
            check(L.dp_halftone(pal.handle, src_ptr, frames, self.h, self.w, self.ht["cell_size"],
                                float(np.cos(a)), float(np.sin(a)), self.ht["dot_gain"],
                                self.ht["min_dot_size"], self.ht["max_dot_size"],
                                HALFTONE_SHAPES.get(self.ht["shape"], 0), self.ht["sharpness"],
                                self.ht_screen.ptr if self.ht_screen is not None else None,
                                dst_ptr, idx_ptr, stream), "dp_halftone")
        elif self.mode == "error_diffusion":
            check(L.dp_error_diffusion(pal.handle, src_ptr, frames, self.h, self.w, self.variant,
                                       int(self.serpentine), dst_ptr, idx_ptr, stream),
                  "dp_error_diffusion")
        elif self.mode == "adaptive_variance":
            check(L.dp_adaptive_variance(pal.handle, src_ptr, frames, self.h, self.w,
                                         self.var_threshold, self.window_radius, dst_ptr, idx_ptr,
                                         stream), "dp_adaptive_variance")
        elif self.mode == "perceptual":
            check(L.dp_perceptual(pal.handle, src_ptr, frames, self.h, self.w, dst_ptr, idx_ptr,
                                  stream), "dp_perceptual")
        elif self.mode == "hybrid":
            check(L.dp_hybrid(pal.handle, src_ptr, frames, self.h, self.w, self.lum_factor,
                              self.col_factor, dst_ptr, idx_ptr, stream), "dp_hybrid")
        else:
            check(L.dp_ostromoukhov(pal.handle, src_ptr, frames, self.h, self.w,
                                    self.coeffs.ctypes.data, int(self.serpentine), dst_ptr,
                                    idx_ptr, stream), "dp_ostromoukhov")


class ChainPlan:
    """pixelize -> dither -> final resize as separate device passes, for the modes whose kernels run
    at identity geometry (halftone, the diffusion family) or for output sizes that are not an exact
    multiple (odd sizes bumped to even, video_processor.py:412-417).  Same interface as Plan.run;
    the intermediate frames live in device buffers owned by the chain (sized for ``max_frames``)."""

    def __init__(self, mode: str, params: Optional[dict], src_hw: Tuple[int, int],
                 dither_hw: Tuple[int, int], out_hw: Tuple[int, int], max_frames: int):
        self.src_h, self.src_w = src_hw
        self.h, self.w = dither_hw
        self.out_h, self.out_w = out_hw
        self.upscale = 1
        self.max_frames = int(max_frames)
        self.inner = Plan(mode, params, self.h, self.w)
        self.mode = mode
        self.pre = (self.src_h, self.src_w) != (self.h, self.w)
        self.post = (self.out_h, self.out_w) != (self.h, self.w)
        n = max(self.max_frames * self.h * self.w * 3, 16)
        self.small = _acquire(n) if self.pre else None       # from the device buffer cache
        self.dith = _acquire(n) if self.post else None
        self.fused_geometry = False

    def run(self, pal: PaletteHandle, src_ptr: int, frames: int, dst_ptr: Optional[int],
            idx_ptr: Optional[int] = None, stream=None):
        assert frames <= self.max_frames
        cur = src_ptr
        if self.pre:
            resample(src_ptr, frames, self.src_h, self.src_w, self.h, self.w, self.small.ptr, stream)
            cur = self.small.ptr
        if self.post and dst_ptr is not None:
            self.inner.run(pal, cur, frames, self.dith.ptr, idx_ptr, stream)
            resample(self.dith.ptr, frames, self.h, self.w, self.out_h, self.out_w, dst_ptr, stream)
        else:
            self.inner.run(pal, cur, frames, dst_ptr, idx_ptr, stream)

    def close(self):
        """Call after the stream the chain ran on has been synchronised."""
        for b in (self.small, self.dith):
            if b is not None:
                _release(b)
        self.small = self.dith = None


def video_geometry(H: int, W: int, pixelize_max_size: Optional[int], final_multiplier: Optional[int],
                   even_final: bool = True):
    """(dither size, output size) of _process_single_frame (video_processor.py:443-462):
    pixelize_regular to even dimensions, then the integer up-scale with odd sizes bumped to even
    (``even_final``; the CLI image path, dither_cli.py:559-566, does not bump)."""
    if pixelize_max_size:
        w, h = even_dimensions(W, H, int(pixelize_max_size))
    else:
        h, w = H, W
    m = int(final_multiplier) if final_multiplier else 1
    oh, ow = h * m, w * m
    if even_final and final_multiplier:
        oh += oh % 2
        ow += ow % 2
    return (h, w), (oh, ow), m


def make_plan(mode: str, params: Optional[dict], H: int, W: int,
              pixelize_max_size: Optional[int] = None, final_multiplier: Optional[int] = None,
              even_final: bool = True, max_frames: int = 1):
    """The cheapest device plan for (mode, geometry): the fused pixelize -> dither -> up-scale
    kernel when the mode has one and the output is an exact multiple, else a ChainPlan."""
    (h, w), (oh, ow), m = video_geometry(H, W, pixelize_max_size, final_multiplier, even_final)
    probe = Plan(mode, params, h, w)
    if (h, w) == (H, W) and (oh, ow) == (h, w):
        return probe
    if probe.fused_geometry and (oh, ow) == (h * m, w * m):
        return Plan(mode, params, h, w, (H, W), m)
    return ChainPlan(mode, params, (H, W), (h, w), (oh, ow), max_frames)


def resample(src_ptr: int, frames: int, src_h: int, src_w: int, dst_h: int, dst_w: int,
             dst_ptr: int, stream=None):
    yt = device_table(nearest_table(src_h, dst_h))
    xt = device_table(nearest_table(src_w, dst_w))
    check(lib().dp_resample_nearest(src_ptr, frames, src_h, src_w, yt.ptr, xt.ptr, dst_h, dst_w,
                                    dst_ptr, stream), "dp_resample_nearest")


# ----------------------------------------------------------------------------------------
# device buffer cache for the numpy-level entry point: cudaMalloc/cudaFree cost about as much
# as dithering a 1080p frame, and callers (GUI preview, per-frame video loop) come back with
# the same sizes.  Buffers return to the cache only after the call's final synchronise.
# ----------------------------------------------------------------------------------------
_buf_cache: Dict[tuple, list] = {}
_buf_lock = threading.Lock()
_buf_cached_bytes = 0
_BUF_CACHE_MAX = 1 << 30


def _acquire(nbytes: int) -> DeviceBuffer:
    size = max(1 << 20, (int(nbytes) + (1 << 20) - 1) >> 20 << 20)
    key = (_capi.ensure_device(), size)
    global _buf_cached_bytes
    with _buf_lock:
        lst = _buf_cache.get(key)
        if lst:
            _buf_cached_bytes -= size
            return lst.pop()
    return DeviceBuffer(size)


def _release(buf: DeviceBuffer):
    global _buf_cached_bytes
    key = (_capi.ensure_device(), buf.nbytes)
    with _buf_lock:
        if _buf_cached_bytes + buf.nbytes <= _BUF_CACHE_MAX:
            _buf_cache.setdefault(key, []).append(buf)
            _buf_cached_bytes += buf.nbytes
            return
    buf.free()


# ----------------------------------------------------------------------------------------
# numpy-level entry point (H2D + kernels + D2H): what the drop-in classes call
# ----------------------------------------------------------------------------------------

def dither_frames(frames_u8: np.ndarray, palette, mode: str, params: Optional[dict] = None,
                  use_gamma: bool = False, pixelize_max_size: Optional[int] = None,
                  final_multiplier: Optional[int] = None, even_final: bool = True,
                  return_indices: bool = False, search_space: bool = False,
                  indices_only: bool = False):
    """uint8 [F,H,W,3] (or [H,W,3]) -> uint8 [F,H',W',3] through the GPU (simple synchronous
    path: one upload, the kernels, one download; clips go through pipeline.FramePipeline).

    pixelize_max_size: regular pixelization first (video_processor.py:563-577).
    final_multiplier:  integer up-scale afterwards (video_processor.py:393-420 when
                       ``even_final`` else dither_cli.py:559-566).
    return_indices:    also return the palette-index plane u8 [F,h,w] of the dithered image.
    indices_only:      return ONLY the index plane (the colour bytes are never produced).
    """
    arr = np.ascontiguousarray(frames_u8, dtype=np.uint8)
    single = arr.ndim == 3
    if single:
        arr = arr[None]
    F, H, W, _ = arr.shape
    pal = get_palette(palette, use_gamma, search_space)
    (h, w), (oh, ow), m = video_geometry(H, W, pixelize_max_size, final_multiplier, even_final)
    want_idx = return_indices or indices_only
    plan = make_plan(mode, params, H, W, pixelize_max_size, final_multiplier, even_final, max_frames=F)
    bufs = []
    try:
        src = _acquire(max(arr.nbytes, 4)).upload(arr)
        bufs.append(src)
        dst = idx_buf = None
        if not indices_only:
            dst = _acquire(max(F * oh * ow * 3, 4))
            bufs.append(dst)
        if want_idx:
            idx_buf = _acquire(max(F * h * w, 4))
            bufs.append(idx_buf)
        plan.run(pal, src.ptr, F, dst.ptr if dst else None, idx_buf.ptr if idx_buf else None)
        out = idx = None
        if dst is not None:
            out = np.empty((F, oh, ow, 3), np.uint8)
            dst.download(out)
        if want_idx:
            idx = np.empty((F, h, w), np.uint8)
            idx_buf.download(idx)
        _capi.sync()
    except BaseException:
        for b in bufs:      # state unknown (a kernel may still run): do not recycle
            b.free()
        raise
    finally:
        if hasattr(plan, "close"):
            _capi.sync()
            plan.close()
    for b in bufs:
        _release(b)
    if single:
        out = out[0] if out is not None else None
        idx = idx[0] if idx is not None else None
    if indices_only:
        return idx
    return (out, idx) if return_indices else out
