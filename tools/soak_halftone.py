"""Soak test: halftone at full size over cell sizes / angles / shapes / gains against the oracle.
    gpurun -- 'python tools/soak_halftone.py'"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import engine, synth  # noqa: E402
from oracle import dither_oracle as O  # noqa: E402  (checker)

t0 = time.time()
bad = n = 0
pals = {"pico8": synth.hex_palette(synth.PICO8), "gb4": synth.hex_palette(synth.GB_POCKET),
        "r64": synth.random_palette(64)}
cases = [{}, {"cell_size": 3, "angle": 90.0}, {"cell_size": 16, "angle": 15.0, "shape": "square"},
         {"cell_size": 5, "angle": 30.0, "shape": "diamond", "sharpness": 1.0},
         {"cell_size": 8, "angle": 45.0, "dot_gain": 1.3, "min_dot_size": 0.1, "max_dot_size": 0.9},
         {"cell_size": 1, "angle": 0.0}, {"cell_size": 64, "angle": 75.0}]
for pname, pal in pals.items():
    for (h, w, nf) in ((1080, 1920, 2), (321, 487, 3), (16, 16, 4)):
        frames = np.stack([synth.frame(h, w, 700 + t) if t % 2 == 0 else synth.noise_frame(h, w, 800 + t)
                           for t in range(nf)])
        for params in cases:
            out = engine.dither_frames(frames, pal, "halftone", params)
            for t in range(nf):
                ref = O.apply_dithering(frames[t], pal, "halftone", params)
                n += 1
                if not np.array_equal(out[t], ref):
                    bad += 1
                    print("MISMATCH", pname, params, h, w, t, int((out[t] != ref).any(axis=2).sum()))
print(f"soak: {n} frames checked, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
