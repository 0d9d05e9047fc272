"""Tiny invocation of every kernel family (ragged sizes, both palette classes, gamma, fused
geometry, k-means).  Written for compute-sanitizer --tool memcheck; that tool is closed on the
build pool, so it serves as a quick plain run:  gpurun -- 'python tools/sanitize_small.py'"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import engine, kmeans, synth  # noqa: E402

pal16 = synth.hex_palette(synth.PICO8)
pal40 = synth.random_palette(40)
frames = np.stack([synth.noise_frame(40, 48, t) for t in range(3)])
odd = np.stack([synth.frame(21, 37, t) for t in range(2)])
for pal in (pal16, pal40):
    for mode, params in (("none", {}), ("bayer", {"size": "8x8"}), ("IGN", {}), ("polka_dot", {"tile_size": 6}),
                         ("halftone", {}), ("error_diffusion", {"variant": "floyd_steinberg"}),
                         ("error_diffusion", {"variant": "jjn"}), ("error_diffusion", {"variant": "atkinson"}),
                         ("error_diffusion", {"variant": "sierra", "serpentine": "true"}), ("ostromoukhov", {})):
        for arr in (frames, odd):
            out, idx = engine.dither_frames(arr, pal, mode, params, return_indices=True)
            assert out.shape == arr.shape
    out = engine.dither_frames(frames, pal, "blue_noise", {"size": 32}, pixelize_max_size=20, final_multiplier=4)
    out = engine.dither_frames(frames, pal, "bayer", {}, use_gamma=True)
c, it = kmeans.kmeans_fit(frames.reshape(-1, 3)[:3000], 8, 42)
print("sanitize_small ok", c.shape, it)
