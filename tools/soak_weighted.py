"""Soak test of the widened diffusion modes (hybrid, perceptual, adaptive variance) against the
oracle: many frames, three palette sizes incl. tie-heavy lattices, gamma, parameter sweeps;
unclamped work values far outside the colour cube (two-colour palette).
    gpurun -- 'python tools/soak_weighted.py'"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import engine, synth  # noqa: E402
from oracle import dither_oracle as O  # noqa: E402  (checker)

t0 = time.time()
bad = n = 0
pals = {"r16": synth.random_palette(16), "r64": synth.random_palette(64), "r256": synth.random_palette(256),
        "lat27": synth.lattice_palette(27, 1, 127), "lat64": synth.lattice_palette(64, 3, 51),
        "two": np.array([[0, 0, 0], [255, 255, 255]]), "gb4": synth.hex_palette(synth.GB_POCKET)}
jobs = [("hybrid", {}), ("hybrid", {"lum_factor": 0.3, "col_factor": 1.7}), ("perceptual", {}),
        ("adaptive_variance", {}), ("adaptive_variance", {"var_threshold": 40.0, "window_radius": 3}),
        ("adaptive_variance", {"var_threshold": 0.0, "window_radius": 5})]
for pname, pal in pals.items():
    for mode, params in jobs:
        for (h, w, nf) in ((270, 480, 12), (97, 333, 20), (33, 1000, 8)):
            frames = np.stack([synth.frame(h, w, 3000 + 7 * t) if t % 3 == 0 else
                               synth.noise_frame(h, w, 4000 + t) if t % 3 == 1 else
                               synth.blocks_frame(h, w, 5000 + t, 8, 6) for t in range(nf)])
            for gamma in ((False, True) if pname in ("r16", "lat27") and h == 97 else (False,)):
                out = engine.dither_frames(frames, pal, mode, params, use_gamma=gamma)
                for t in range(0, nf, 4):
                    ref = O.apply_dithering(frames[t], pal, mode, params, gamma)
                    n += 1
                    if not np.array_equal(out[t], ref):
                        bad += 1
                        print("MISMATCH", pname, mode, params, h, w, t, gamma, int((out[t] != ref).any(axis=2).sum()))
print(f"soak: {n} frames checked, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
