"""dither_pie_b200 -- B200-native (sm_100a) implementation of dither_pie's per-pixel hot path.

Drop-in modules: ``dither_pie_b200.dithering_lib`` and ``dither_pie_b200.video_processor`` mirror
the reference modules of the same names for the in-scope classes and functions; the per-pixel
work runs in ``libditherpie_b200.so`` (C ABI in include/ditherpie_b200.h).  Importing the
package does not touch CUDA; the first call that needs the GPU loads the library and raises if
it, or a B200, is missing.  There is no CPU fallback.
"""
from .dithering_lib import (  # noqa: F401
    AdaptiveVarianceDitherStrategy, BaseDitherStrategy, BayerDitherStrategy, BlueNoiseDitherStrategy, ColorReducer, DitherMode,
    DitherUtils, ErrorDiffusionDitherStrategy, ErrorDiffusionKernel, HalftoneDitherStrategy,
    HybridDitherStrategy, ImageDitherer, InterleavedGradientNoiseDitherStrategy, MatrixDitherStrategy,
    NoDitherStrategy, OstromoukhovDitherStrategy, PaletteSource, PerceptualDitherStrategy,
    PixelizeMethod, PolkaDotDitherStrategy, generate_blue_noise)
from .video_processor import NeuralPixelizer, VideoProcessor, pixelize_regular, shard_frames  # noqa: F401

__version__ = "0.1.0"
