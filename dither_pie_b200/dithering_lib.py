"""Drop-in surface of the reference's ``dithering_lib`` for the per-pixel hot path.

Same names, constructor arguments, defaults and error behaviour as
dobrosketchkun/dither_pie ``dithering_lib.py`` (``__all__`` at :27-57) for every in-scope class;
the per-pixel work runs in libditherpie_b200.so (hand-written CUDA, sm_100a) through ctypes.
There is no CPU fallback: without the library or without a B200 every ``dither`` call raises.

Out-of-scope modes (riemersma, wavelet -- SURVEY.md section 2, rows 14-15) keep their names so
that imports do not break, and raise NotImplementedError when used.  ``hybrid``, ``perceptual``
and ``adaptive_variance`` run on the error-diffusion wavefront:
hybrid with the semantics of the reference's numba kernel, the other two with those of their
pure-Python loops (f32, KD-tree nearest of the unclamped value).
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import engine

__all__ = [
    'DitherMode', 'PixelizeMethod', 'PaletteSource',
    'ImageDitherer', 'ColorReducer', 'DitherUtils', 'BaseDitherStrategy', 'ErrorDiffusionKernel',
    'NoDitherStrategy', 'MatrixDitherStrategy', 'BayerDitherStrategy', 'BlueNoiseDitherStrategy',
    'InterleavedGradientNoiseDitherStrategy', 'ErrorDiffusionDitherStrategy',
    'OstromoukhovDitherStrategy', 'RiemersmaDitherStrategy', 'PolkaDotDitherStrategy',
    'WaveletDitherStrategy', 'AdaptiveVarianceDitherStrategy', 'PerceptualDitherStrategy',
    'HybridDitherStrategy', 'HalftoneDitherStrategy', 'generate_blue_noise',
]


class DitherMode(Enum):
    """dithering_lib.py:61-75."""
    NONE = "none"
    BAYER = "bayer"
    ERROR_DIFFUSION = "error_diffusion"
    RIEMERSMA = "riemersma"
    BLUE_NOISE = "blue_noise"
    INTERLEAVED_GRADIENT_NOISE = "IGN"
    POLKA_DOT = "polka_dot"
    WAVELET = "wavelet"
    ADAPTIVE_VARIANCE = "adaptive_variance"
    PERCEPTUAL = "perceptual"
    HYBRID = "hybrid"
    HALFTONE = "halftone"
    OSTROMOUKHOV = "ostromoukhov"


class PixelizeMethod(Enum):
    """dithering_lib.py:78-82."""
    NONE = "none"
    REGULAR = "regular"
    NEURAL = "neural"


class PaletteSource(Enum):
    """dithering_lib.py:85-91."""
    MEDIAN_CUT = "median_cut"
    KMEANS = "kmeans"
    UNIFORM = "uniform"
    CUSTOM = "custom"
    FROM_FILE = "file"


class ErrorDiffusionKernel:
    """Tap tables of the eight fixed-weight kernels (dithering_lib.py:96-209).  The CUDA library
    carries its own compile-time copy (csrc/dp_diffusion.cu); this class is the Python-visible
    description, same dict layout as the reference."""

    @staticmethod
    def _k(taps, divisor, description, rows):
        return {'weights': taps, 'divisor': divisor, 'description': description, 'rows': rows}

    _row1_wide = lambda a, b, c: [(-2, 1, a), (-1, 1, b), (0, 1, c), (1, 1, b), (2, 1, a)]  # noqa: E731
    _row2_wide = lambda a, b, c: [(-2, 2, a), (-1, 2, b), (0, 2, c), (1, 2, b), (2, 2, a)]  # noqa: E731

    FLOYD_STEINBERG = _k.__func__([(1, 0, 7), (-1, 1, 3), (0, 1, 5), (1, 1, 1)], 16,
                                  'Classic Floyd-Steinberg (4 neighbors)', 2)
    JJN = _k.__func__([(1, 0, 7), (2, 0, 5)] + _row1_wide(3, 5, 7) + _row2_wide(1, 3, 5), 48,
                      'Jarvis-Judice-Ninke (12 neighbors, smooth gradients)', 3)
    STUCKI = _k.__func__([(1, 0, 8), (2, 0, 4)] + _row1_wide(2, 4, 8) + _row2_wide(1, 2, 4), 42,
                         'Stucki (12 neighbors, photographic quality)', 3)
    BURKES = _k.__func__([(1, 0, 8), (2, 0, 4)] + _row1_wide(2, 4, 8), 32,
                         'Burkes (7 neighbors, fast)', 2)
    ATKINSON = _k.__func__([(1, 0, 1), (2, 0, 1), (-1, 1, 1), (0, 1, 1), (1, 1, 1), (0, 2, 1)], 8,
                           'Atkinson (6 neighbors, classic Mac look)', 3)
    SIERRA = _k.__func__([(1, 0, 5), (2, 0, 3)] + _row1_wide(2, 4, 5)
                         + [(-1, 2, 2), (0, 2, 3), (1, 2, 2)], 32,
                         'Sierra Full (10 neighbors, high quality)', 3)
    SIERRA_TWO_ROW = _k.__func__([(1, 0, 4), (2, 0, 3)] + _row1_wide(1, 2, 3), 16,
                                 'Sierra Two-Row (8 neighbors, balanced)', 2)
    SIERRA_LITE = _k.__func__([(1, 0, 2), (-1, 1, 1), (0, 1, 1)], 4,
                              'Sierra Lite (4 neighbors, fastest)', 2)

    @classmethod
    def list_kernels(cls) -> List[str]:
        return list(engine.ED_VARIANTS)

    @classmethod
    def get_kernel(cls, name: str) -> Dict[str, Any]:
        """Unknown names fall back to Floyd-Steinberg (:203)."""
        return getattr(cls, name.upper(), cls.FLOYD_STEINBERG) if name in engine.ED_VARIANTS \
            else cls.FLOYD_STEINBERG


# ----------------------------------------------------------------------------------------
# strategies
# ----------------------------------------------------------------------------------------

def _pixels_to_u8(pixels: np.ndarray, image_size: Tuple[int, int]) -> np.ndarray:
    """The strategy API hands over f32 [N,3] (dithering_lib.py:1977); on the B200 path pixels
    are bytes.  ImageDitherer always passes integral values; anything else is refused loudly."""
    h, w = image_size
    px = np.asarray(pixels)
    if px.shape != (h * w, 3):
        raise ValueError(f"pixels must have shape ({h * w}, 3), got {px.shape}")
    u8 = px.astype(np.uint8)
    if not np.array_equal(u8, px):
        raise ValueError("dither_pie_b200 strategies take byte-valued pixels (0..255 integers), "
                         "as ImageDitherer.apply_dithering always passes")
    return u8.reshape(h, w, 3)


class BaseDitherStrategy:
    """dithering_lib.py:313-330.  ``dither(pixels f32[N,3], palette_arr f32[K,3], (h,w))``
    returns palette rows, shape [N,3]."""
    _mode: Optional[str] = None

    def dither(self, pixels: np.ndarray, palette_arr: np.ndarray,
               image_size: Tuple[int, int]) -> np.ndarray:
        if self._mode is None:
            raise NotImplementedError
        img = _pixels_to_u8(pixels, image_size)
        pal = np.asarray(palette_arr)
        idx = engine.dither_frames(img, pal, self._mode, self.get_current_parameters(),
                                   use_gamma=False, search_space=True, indices_only=True)
        return pal[idx.reshape(-1).astype(np.int32), :]

    @staticmethod
    def get_parameter_info() -> Optional[Dict[str, Any]]:
        return None

    def get_current_parameters(self) -> Dict[str, Any]:
        return {}


class NoDitherStrategy(BaseDitherStrategy):
    """Nearest palette colour (:333-341)."""
    _mode = "none"


class MatrixDitherStrategy(BaseDitherStrategy):
    """Threshold-matrix dithering (:346-378) with an arbitrary f32 matrix."""
    _mode = "matrix"

    def __init__(self, threshold_matrix: np.ndarray):
        self.threshold_matrix = threshold_matrix

    def dither(self, pixels, palette_arr, image_size):
        img = _pixels_to_u8(pixels, image_size)
        pal = np.asarray(palette_arr)
        idx = _run_matrix(img, pal, np.asarray(self.threshold_matrix, np.float32))
        return pal[idx.reshape(-1).astype(np.int32), :]


def _run_matrix(img_u8: np.ndarray, palette_arr: np.ndarray, matrix: np.ndarray) -> np.ndarray:
    """Custom-matrix path for MatrixDitherStrategy subclasses constructed directly: the index
    plane only, device buffers from the engine's buffer cache."""
    import ctypes as C
    from ._capi import Geometry, check, lib, sync
    h, w, _ = img_u8.shape
    pal = engine.get_palette(palette_arr, False, True)
    geo = Geometry()
    geo.src_h = geo.h = h
    geo.src_w = geo.w = w
    geo.upscale = 1
    mat = engine.device_table(np.ascontiguousarray(matrix, np.float32))
    src = engine._acquire(max(img_u8.nbytes, 4)).upload(np.ascontiguousarray(img_u8))
    ib = engine._acquire(max(h * w, 4))
    try:
        check(lib().dp_threshold_dither(pal.handle, src.ptr, 1, C.byref(geo), 1, mat.ptr,
                                        matrix.shape[0], matrix.shape[1], 0.0, 0.0, 1.0, None,
                                        ib.ptr, None), "dp_threshold_dither")
        idx = np.empty((h, w), np.uint8)
        ib.download(idx)
        sync()
    except BaseException:
        src.free()
        ib.free()
        raise
    engine._release(src)
    engine._release(ib)
    return idx


def generate_blue_noise(size: int = 64, seed: int = 42) -> np.ndarray:
    """Blue-noise threshold matrix, bit-identical to dithering_lib.py:381-399."""
    return engine.blue_noise_matrix(size, seed).copy()


class BayerDitherStrategy(MatrixDitherStrategy):
    """:402-448."""
    _mode = "bayer"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {'size': {'type': 'choice', 'default': '4x4',
                         'choices': ['2x2', '4x4', '8x8', '16x16', 'psx4x4'], 'label': 'Matrix',
                         'description': 'Bayer matrix size or PSX 4x4 variant '
                                        '(larger = finer patterns)'}}

    def __init__(self, size: str = '4x4'):
        self.size = size
        super().__init__(engine.bayer_matrix(size))

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'size': self.size}

    dither = BaseDitherStrategy.dither


class BlueNoiseDitherStrategy(MatrixDitherStrategy):
    """:451-499 (matrices cached per (size, seed) like the reference's class-level cache)."""
    _mode = "blue_noise"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'size': {'type': 'int', 'default': 64, 'min': 32, 'max': 128, 'label': 'Matrix Size',
                     'description': 'Size of the blue noise matrix '
                                    '(larger = more detail but slower)'},
            'seed': {'type': 'int', 'default': 42, 'min': 0, 'max': 9999, 'label': 'Random Seed',
                     'description': 'Seed for noise generation '
                                    '(different seeds = different patterns)'},
        }

    def __init__(self, size: int = 64, seed: int = 42):
        self.size = size
        self.seed = seed
        super().__init__(engine.blue_noise_matrix(size, seed))

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'size': self.size, 'seed': self.seed}

    dither = BaseDitherStrategy.dither


class InterleavedGradientNoiseDitherStrategy(BaseDitherStrategy):
    """:502-571."""
    _mode = "IGN"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'scale': {'type': 'float', 'default': 1.0, 'min': 0.1, 'max': 10.0, 'step': 0.1,
                      'label': 'Scale',
                      'description': 'Noise frequency (lower = larger pattern, '
                                     'higher = finer grain)'},
            'seed': {'type': 'int', 'default': 0, 'min': 0, 'max': 9999, 'label': 'Seed',
                     'description': 'Deterministic offset to shift the pattern'},
        }

    def __init__(self, scale: float = 1.0, seed: int = 0):
        self.scale = float(scale)
        self.seed = int(seed)

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'scale': self.scale, 'seed': self.seed}


class ErrorDiffusionDitherStrategy(BaseDitherStrategy):
    """:576-690 (the numba path's semantics; `serpentine` is the STRING 'true'/'false')."""
    _mode = "error_diffusion"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'variant': {'type': 'choice', 'default': 'atkinson',
                        'choices': ErrorDiffusionKernel.list_kernels(), 'label': 'Algorithm',
                        'description': 'Error diffusion algorithm variant'},
            'serpentine': {'type': 'choice', 'default': 'false', 'choices': ['true', 'false'],
                           'label': 'Serpentine Scan',
                           'description': 'Alternates direction each row to reduce artifacts'},
        }

    def __init__(self, variant: str = 'atkinson', serpentine: str = 'false'):
        self.variant = variant
        self.serpentine = (serpentine == 'true')
        self._kernel = ErrorDiffusionKernel.get_kernel(variant)

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'variant': self.variant, 'serpentine': 'true' if self.serpentine else 'false'}


class PolkaDotDitherStrategy(BaseDitherStrategy):
    """:695-766."""
    _mode = "polka_dot"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'tile_size': {'type': 'int', 'default': 8, 'min': 4, 'max': 32, 'label': 'Tile Size',
                          'description': 'Size of the repeating dot pattern'},
            'gamma': {'type': 'float', 'default': 1.5, 'min': 0.5, 'max': 3.0, 'step': 0.1,
                      'label': 'Gamma',
                      'description': 'Controls dot shape curve (higher = sharper edges)'},
        }

    def __init__(self, tile_size: int = 8, gamma: float = 1.5):
        self.tile_size = tile_size
        self.gamma = gamma
        self.threshold_matrix = engine.polka_dot_matrix(tile_size, gamma)

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'tile_size': self.tile_size, 'gamma': self.gamma}


class OstromoukhovDitherStrategy(BaseDitherStrategy):
    """:1160-1269, the live (pure-Python + KD-tree) path's semantics."""
    _mode = "ostromoukhov"
    COEFFS_TABLE = [tuple(int(v) for v in row) for row in engine.ostromoukhov_coeffs()]

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {'serpentine': {'type': 'choice', 'default': 'false', 'choices': ['true', 'false'],
                               'label': 'Serpentine Scan',
                               'description': 'Alternates direction each row to reduce artifacts'}}

    def __init__(self, serpentine: str = 'false'):
        self.serpentine = (serpentine == 'true')

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'serpentine': 'true' if self.serpentine else 'false'}


class HalftoneDitherStrategy(BaseDitherStrategy):
    """:1498-1695."""
    _mode = "halftone"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'cell_size': {'type': 'int', 'default': 8, 'min': 2, 'max': 32, 'label': 'Cell Size',
                          'description': 'Distance between dot centers (smaller = finer detail)'},
            'angle': {'type': 'float', 'default': 45.0, 'min': 0.0, 'max': 90.0,
                      'label': 'Screen Angle',
                      'description': 'Rotation angle in degrees (45° is classic newspaper)'},
            'dot_gain': {'type': 'float', 'default': 1.0, 'min': 0.5, 'max': 3.0, 'step': 0.1,
                         'label': 'Dot Gain',
                         'description': 'Controls dot growth (1.0 = linear, '
                                        'higher = more contrast)'},
            'min_dot_size': {'type': 'float', 'default': 0.0, 'min': 0.0, 'max': 0.5,
                             'step': 0.05, 'label': 'Min Dot Size',
                             'description': 'Minimum dot threshold (0 = pure white possible)'},
            'max_dot_size': {'type': 'float', 'default': 1.0, 'min': 0.5, 'max': 1.0,
                             'step': 0.05, 'label': 'Max Dot Size',
                             'description': 'Maximum dot threshold (1.0 = pure black possible)'},
            'shape': {'type': 'choice', 'default': 'circle',
                      'choices': ['circle', 'square', 'diamond'], 'label': 'Dot Shape',
                      'description': 'Shape of halftone dots'},
            'sharpness': {'type': 'float', 'default': 1.5, 'min': 0.5, 'max': 4.0, 'step': 0.1,
                          'label': 'Sharpness',
                          'description': 'Edge sharpness (higher = crisper dots)'},
        }

    def __init__(self, cell_size: int = 8, angle: float = 45.0, dot_gain: float = 1.0,
                 min_dot_size: float = 0.0, max_dot_size: float = 1.0, shape: str = "circle",
                 sharpness: float = 1.5):
        self.cell_size = cell_size
        self.angle = angle
        self.dot_gain = dot_gain
        self.min_dot_size = min_dot_size
        self.max_dot_size = max_dot_size
        self.shape = shape
        self.sharpness = sharpness

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'cell_size': self.cell_size, 'angle': self.angle, 'dot_gain': self.dot_gain,
                'min_dot_size': self.min_dot_size, 'max_dot_size': self.max_dot_size,
                'shape': self.shape, 'sharpness': self.sharpness}


def _out_of_scope(name: str, where: str):
    class _Stub(BaseDitherStrategy):
        __doc__ = (f"{name}: outside the B200 hot path (reference {where}; SURVEY.md section 2). "
                   "Kept so that imports resolve; using it raises.")

        def __init__(self, *a, **k):
            raise NotImplementedError(
                f"{name} is outside the per-pixel hot path built for B200 "
                f"(reference dithering_lib.py {where}); use the reference for this mode")

    _Stub.__name__ = _Stub.__qualname__ = name
    return _Stub


RiemersmaDitherStrategy = _out_of_scope('RiemersmaDitherStrategy', ':771-841')
WaveletDitherStrategy = _out_of_scope('WaveletDitherStrategy', ':846-941')


class AdaptiveVarianceDitherStrategy(BaseDitherStrategy):
    """:946-1025: Floyd-Steinberg that distributes a pixel's error only where the local variance
    of the gray image reaches ``var_threshold``."""
    _mode = "adaptive_variance"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'var_threshold': {'type': 'float', 'default': 300.0, 'min': 0.0, 'max': 1000.0,
                              'step': 10.0, 'label': 'Variance Threshold',
                              'description': 'Threshold for local variance to trigger error diffusion'},
            'window_radius': {'type': 'int', 'default': 1, 'min': 1, 'max': 5,
                              'label': 'Window Radius',
                              'description': 'Radius of window for computing local variance'},
        }

    def __init__(self, var_threshold: float = 300.0, window_radius: int = 1):
        self.var_threshold = var_threshold
        self.window_radius = window_radius

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'var_threshold': self.var_threshold, 'window_radius': self.window_radius}


class PerceptualDitherStrategy(BaseDitherStrategy):
    """:1030-1066: Floyd-Steinberg whose taps are scaled by a luminance factor of the original
    pixel.  Only the default ``base_weights`` run on the GPU; a custom list raises."""
    _mode = "perceptual"
    _FS = [(1, 0, 7 / 16), (-1, 1, 3 / 16), (0, 1, 5 / 16), (1, 1, 1 / 16)]

    def __init__(self, base_weights=None):
        if base_weights is not None and [tuple(t) for t in base_weights] != self._FS:
            raise NotImplementedError(
                "PerceptualDitherStrategy: only the default Floyd-Steinberg base_weights are "
                "built for B200; use the reference for custom weights")
        self.base_weights = list(self._FS)


class HybridDitherStrategy(BaseDitherStrategy):
    """:1071-1155 with the semantics of its numba core ``_hybrid_numba`` (:1396-1494), the path
    the reference takes whenever numba imports: Floyd-Steinberg diffusion of
    ``lum_factor * luminance part + col_factor * colour part`` of the error."""
    _mode = "hybrid"

    @staticmethod
    def get_parameter_info() -> Dict[str, Any]:
        return {
            'lum_factor': {'type': 'float', 'default': 1.0, 'min': 0.0, 'max': 2.0, 'step': 0.1,
                           'label': 'Luminance Factor',
                           'description': 'Strength of luminance error diffusion '
                                          '(1.0 = full, 0.0 = none)'},
            'col_factor': {'type': 'float', 'default': 0.2, 'min': 0.0, 'max': 2.0, 'step': 0.1,
                           'label': 'Color Factor',
                           'description': 'Strength of color error diffusion '
                                          '(lower = less color noise)'},
        }

    def __init__(self, lum_factor: float = 1.0, col_factor: float = 0.2):
        self.lum_factor = lum_factor
        self.col_factor = col_factor
        self.fs_offsets = [(1, 0, 7 / 16), (-1, 1, 3 / 16), (0, 1, 5 / 16), (1, 1, 1 / 16)]

    def get_current_parameters(self) -> Dict[str, Any]:
        return {'lum_factor': self.lum_factor, 'col_factor': self.col_factor}


# ----------------------------------------------------------------------------------------
# utils, colour reduction, image wrapper
# ----------------------------------------------------------------------------------------

class DitherUtils:
    """:1700-1802."""
    BAYER2x2 = engine.bayer_matrix('2x2')
    BAYER4x4 = engine.bayer_matrix('4x4')
    BAYER8x8 = engine.bayer_matrix('8x8')
    BAYER16x16 = engine.bayer_matrix('16x16')
    PSX4x4 = engine.bayer_matrix('psx4x4')

    @staticmethod
    def get_threshold_matrix(mode: DitherMode, size: str = '4x4') -> np.ndarray:
        if mode == DitherMode.NONE:
            return np.ones((1, 1), dtype=np.float32)
        if mode == DitherMode.BAYER:
            return engine.bayer_matrix(size)
        raise ValueError(f"Unsupported matrix mode: {mode}")

    srgb_to_linear = staticmethod(engine.srgb_to_linear)
    linear_to_srgb = staticmethod(engine.linear_to_srgb)


class ColorReducer:
    """:1807-1872.  k-means runs on the GPU; median cut and the uniform cube are palette
    SET-UP on the host ('next' #1), restated here so that
    ``ImageDitherer(palette=None)`` behaves like the reference."""

    @staticmethod
    def find_dominant_channel(colors) -> int:
        spans = [max(c[ch] for c in colors) - min(c[ch] for c in colors) for ch in range(3)]
        return spans.index(max(spans))

    @staticmethod
    def median_cut(colors, depth: int):
        if depth == 0 or len(colors) == 0:
            if not colors:
                return [(0, 0, 0)]
            return [tuple(int(sum(ch) / len(ch)) for ch in zip(*colors))]
        ch = ColorReducer.find_dominant_channel(colors)
        colors.sort(key=lambda c: c[ch])
        mid = len(colors) // 2
        return (ColorReducer.median_cut(colors[:mid], depth - 1)
                + ColorReducer.median_cut(colors[mid:], depth - 1))

    @staticmethod
    def _median_cut_array(cols: np.ndarray, depth: int):
        """median_cut on a PLANAR uint8 [3,N] array whose column order is the list order of the
        reference: ``list.sort(key=channel)`` is a stable sort, so a stable argsort on the
        channel reproduces it (numpy sorts bytes with a radix sort); the box average is the same
        float division of exact integer sums, truncated."""
        n = cols.shape[1]
        if depth == 0 or n == 0:
            if n == 0:
                return [(0, 0, 0)]
            return [tuple(int(int(sv) / n) for sv in cols.sum(axis=1, dtype=np.int64))]
        spans = [int(cols[c].max()) - int(cols[c].min()) for c in range(3)]
        ch = spans.index(max(spans))
        cols = np.take(cols, np.argsort(cols[ch], kind='stable'), axis=1)
        mid = n // 2
        return (ColorReducer._median_cut_array(cols[:, :mid], depth - 1)
                + ColorReducer._median_cut_array(cols[:, mid:], depth - 1))

    @staticmethod
    def unique_colors_in_set_order(arr_u8: np.ndarray) -> np.ndarray:
        """``list(set(image.getdata()))`` (:1837) as a uint8 [N,3] array: the unique colours in
        the iteration order of a CPython set of (r, g, b) tuples, replayed natively
        (csrc/dp_pyset.cu; host code, no GPU needed).  That order decides how equal keys fall
        around each median (``list.sort`` is stable)."""
        from . import _capi
        import ctypes as C
        flat = np.ascontiguousarray(arr_u8, np.uint8).reshape(-1, 3)
        out = np.empty_like(flat)
        n = C.c_int64(0)
        _capi.check(_capi.lib().dp_unique_colors_pyset_order(
            flat.ctypes.data, flat.shape[0], out.ctypes.data, C.byref(n)),
            "dp_unique_colors_pyset_order")
        return out[:n.value]

    @staticmethod
    def reduce_colors(image, num_colors: int):
        """:1834-1843.  The unique colours enter the cut in CPython set order
        (unique_colors_in_set_order); the recursive sort/split/average work -- most of the
        reference's 7.6 s on a 1080p frame -- runs on arrays."""
        image = image.convert('RGB')
        arr = ColorReducer.unique_colors_in_set_order(np.asarray(image, dtype=np.uint8))
        num_colors = max(1, num_colors)
        depth = int(math.log2(num_colors)) if num_colors > 1 else 0
        if arr.shape[0] < 64:
            return ColorReducer.median_cut([tuple(int(v) for v in c) for c in arr], depth)
        return ColorReducer._median_cut_array(np.ascontiguousarray(arr.T), depth)

    @staticmethod
    def generate_kmeans_palette(img, num_colors: int, random_state=42):
        from .kmeans import kmeans_palette
        arr = np.array(img.convert('RGB'))
        return kmeans_palette(arr, num_colors, random_state)

    @staticmethod
    def generate_uniform_palette(num_colors: int):
        cube = int(math.ceil(num_colors ** (1 / 3)))
        out = []
        for r in range(cube):
            for g in range(cube):
                for b in range(cube):
                    if len(out) >= num_colors:
                        break
                    out.append(tuple(int(v * 255 / (cube - 1)) if cube > 1 else 128
                                     for v in (r, g, b)))
        return out[:num_colors]


_STRATEGIES = {
    DitherMode.NONE: NoDitherStrategy,
    DitherMode.BAYER: BayerDitherStrategy,
    DitherMode.BLUE_NOISE: BlueNoiseDitherStrategy,
    DitherMode.INTERLEAVED_GRADIENT_NOISE: InterleavedGradientNoiseDitherStrategy,
    DitherMode.POLKA_DOT: PolkaDotDitherStrategy,
    DitherMode.ERROR_DIFFUSION: ErrorDiffusionDitherStrategy,
    DitherMode.RIEMERSMA: RiemersmaDitherStrategy,
    DitherMode.WAVELET: WaveletDitherStrategy,
    DitherMode.ADAPTIVE_VARIANCE: AdaptiveVarianceDitherStrategy,
    DitherMode.PERCEPTUAL: PerceptualDitherStrategy,
    DitherMode.HYBRID: HybridDitherStrategy,
    DitherMode.HALFTONE: HalftoneDitherStrategy,
    DitherMode.OSTROMOUKHOV: OstromoukhovDitherStrategy,
}


class ImageDitherer:
    """:1877-1992.  Plain attributes only, so instances stay picklable like the reference's
    (video_processor.py:312-322 pickles the ditherer into pool workers)."""

    def __init__(self, num_colors: int = 16, dither_mode: Optional[DitherMode] = DitherMode.BAYER,
                 palette: Optional[List[Tuple[int, int, int]]] = None, use_gamma: bool = False,
                 dither_params: Optional[Dict[str, Any]] = None):
        self.num_colors = num_colors
        self.dither_mode = dither_mode
        self.palette = palette
        self.use_gamma = use_gamma
        self.dither_params = dither_params or {}

    @staticmethod
    def get_mode_parameters(mode: DitherMode) -> Optional[Dict[str, Any]]:
        cls = _STRATEGIES.get(mode)
        if cls is None or cls in (NoDitherStrategy,):
            return None
        if getattr(cls, '_mode', None) is None and not hasattr(cls, 'get_parameter_info'):
            return None
        try:
            return cls.get_parameter_info()
        except Exception:
            return None

    @staticmethod
    def mode_has_parameters(mode: DitherMode) -> bool:
        return ImageDitherer.get_mode_parameters(mode) is not None

    def _get_dither_strategy(self, mode: DitherMode) -> BaseDitherStrategy:
        """:1918-1950: defaults + user overrides -> constructor kwargs (unknown keys raise
        TypeError from the constructor, as in the reference)."""
        cls = _STRATEGIES.get(mode)
        if cls is None:
            raise ValueError(f"Unrecognized DitherMode: {mode}")
        info = cls.get_parameter_info()
        if info:
            settings = {k: v['default'] for k, v in info.items()}
            settings.update(self.dither_params)
            return cls(**settings)
        return cls()

    def _ensure_palette(self, arr_u8: np.ndarray):
        """palette=None -> median cut, stored on the instance (:1960-1966)."""
        if self.palette is None:
            from PIL import Image
            src = arr_u8
            if self.use_gamma:
                src = engine.gamma_in_lut()[arr_u8]
            self.palette = ColorReducer.reduce_colors(Image.fromarray(src, 'RGB'), self.num_colors)

    def apply_dithering_array(self, arr_u8: np.ndarray) -> np.ndarray:
        """uint8 [h,w,3] or [F,h,w,3] -> uint8, same shape.  The array-level twin of
        apply_dithering (no PIL round trip); frames of a batch share the palette."""
        arr_u8 = np.ascontiguousarray(arr_u8, np.uint8)
        self._ensure_palette(arr_u8 if arr_u8.ndim == 3 else arr_u8[0])
        if not self.dither_mode:
            self.dither_mode = DitherMode.NONE
        strategy = self._get_dither_strategy(self.dither_mode)  # validates mode + kwargs
        return engine.dither_frames(arr_u8, self.palette, strategy._mode,
                                    strategy.get_current_parameters(), use_gamma=self.use_gamma)

    def apply_dithering(self, image):
        """PIL.Image -> PIL.Image 'RGB' (:1952-1992)."""
        from PIL import Image
        arr = np.array(image.convert('RGB'), dtype=np.uint8)
        return Image.fromarray(self.apply_dithering_array(arr), 'RGB')
