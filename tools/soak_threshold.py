"""Soak test: full-size frames of the threshold family against the oracle (v4 tiles straddling
frames, ragged tails, tie table, sub-cell refinement, fused geometry).
    gpurun -- 'python tools/soak_threshold.py'"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import engine, synth  # noqa: E402
from oracle import dither_oracle as O  # noqa: E402  (checker)

t0 = time.time()
bad = n = 0
modes = [("none", {}), ("bayer", {"size": "8x8"}), ("bayer", {"size": "16x16"}), ("IGN", {"scale": 2.5, "seed": 17}),
         ("blue_noise", {"size": 32, "seed": 5}), ("polka_dot", {"tile_size": 6})]
pals = {"pico8": synth.hex_palette(synth.PICO8), "c64": synth.hex_palette(synth.C64),
        "r30": synth.random_palette(30, seed=9), "r31": synth.random_palette(31, seed=9),
        "lat27": synth.lattice_palette(27, 1, 127),
        "r100": synth.random_palette(100, seed=9), "r256": synth.random_palette(256)}   # wide format: deferred fixes
for pname, pal in pals.items():
    for (h, w, nf) in ((1080, 1920, 3), (720, 1296, 5), (33, 48, 7)):
        frames = np.stack([synth.noise_frame(h, w, 300 + t) if t % 2 else synth.frame(h, w, 400 + t)
                           for t in range(nf)])
        for mode, params in modes:
            out = engine.dither_frames(frames, pal, mode, params)
            for t in (0, nf - 1):
                ref = O.apply_dithering(frames[t], pal, mode, params)
                n += 1
                if not np.array_equal(out[t], ref):
                    bad += 1
                    print("MISMATCH", pname, mode, params, h, w, t, int((out[t] != ref).any(axis=2).sum()))
    # fused pixelise -> dither -> x3 against the composition of the oracle's pieces
    frames = np.stack([synth.frame(1080, 1920, 500 + t) for t in range(2)])
    out = engine.dither_frames(frames, pal, "bayer", {"size": "4x4"}, pixelize_max_size=270, final_multiplier=3)
    for t in range(2):
        small = O.pixelize_regular(frames[t], 270)
        ref = O.final_resize(O.apply_dithering(small, pal, "bayer", {"size": "4x4"}), 3)
        n += 1
        if not np.array_equal(out[t], ref):
            bad += 1
            print("MISMATCH fused", pname, t)
print(f"soak: {n} frames checked, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
