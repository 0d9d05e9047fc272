"""Soak test: many frames / variants / palette sizes of the wavefront kernel against the oracle
(the ready queue, the hand-off flags and the chunked writeback under real concurrency).
    gpurun -- 'python tools/soak_diffusion.py'"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import engine, synth  # noqa: E402
from oracle import dither_oracle as O  # noqa: E402  (checker)

t0 = time.time()
bad = 0
n = 0
for K in (16, 64, 256):
    pal = synth.random_palette(K)
    for variant in list(O.ED_KERNELS) + ["ostromoukhov"]:
        for (h, w, nf) in ((270, 480, 24), (540, 960, 12), (97, 333, 40)):
            frames = np.stack([synth.frame(h, w, 1000 + 7 * t + K) if t % 3 else synth.noise_frame(h, w, 2000 + t)
                               for t in range(nf)])
            if variant == "ostromoukhov":
                if (h, w) == (540, 960):
                    continue   # the oracle's serial loop is slow for this one
                mode, params = "ostromoukhov", {}
            else:
                mode, params = "error_diffusion", {"variant": variant}
            out = engine.dither_frames(frames, pal, mode, params)
            for t in range(0, nf, 5):
                ref = O.apply_dithering(frames[t], pal, mode, params)
                n += 1
                if not np.array_equal(out[t], ref):
                    bad += 1
                    print("MISMATCH", K, variant, h, w, t, int((out[t] != ref).any(axis=2).sum()))
print(f"soak: {n} frames checked, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
