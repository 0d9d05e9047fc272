"""Soak test + timing of the serial (serpentine) diffusion kernel against the oracle.
    gpurun -- python tools/soak_serpentine.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import engine, synth
from oracle import dither_oracle as O
bad = 0
# h < 2 images and weighted/hybrid modes also take the serial kernel when serpentine is on
for (h, w) in ((1, 50), (37, 61), (96, 160)):
    for K, pal in ((16, synth.hex_palette(synth.PICO8)), (256, synth.random_palette(256)), (27, synth.lattice_palette(27, 1, 127))):
        img = synth.frame(h, w, 3) if h > 1 else synth.frame(2, w, 3)[:1]
        blk = synth.blocks_frame(h, w, 9, 4, 6) if h > 8 else img
        for im in (img, blk):
            for mode, params in [("error_diffusion", {"variant": v, "serpentine": "true"}) for v in
                                 ("floyd_steinberg", "jjn", "stucki", "burkes", "atkinson", "sierra", "sierra_two_row", "sierra_lite")] + \
                                [("ostromoukhov", {"serpentine": "true"})]:
                out = engine.dither_frames(im, pal, mode, params)
                ref = O.apply_dithering(im, pal, mode, params)
                nb = int((out != ref).any(axis=2).sum())
                if nb:
                    bad += 1
                    print("MISMATCH", h, w, K, mode, params, nb)
print("serpentine check: mismatching cases", bad)
pal = synth.random_palette(256)
for (h, w) in ((1080, 1920), (2160, 3840)):
    img = synth.frame(h, w, 1)
    for mode, params in (("error_diffusion", {"variant": "floyd_steinberg", "serpentine": "true"}), ("ostromoukhov", {"serpentine": "true"})):
        engine.dither_frames(img[:64], pal, mode, params)
        t0 = time.perf_counter(); out = engine.dither_frames(img, pal, mode, params); dt = time.perf_counter() - t0
        print(f"{h}x{w} {mode} serpentine K=256: {dt*1e3:.1f} ms  {h*w/dt/1e6:.2f} Mpx/s")
        if h == 1080:
            t0 = time.perf_counter(); ref = O.apply_dithering(img, pal, mode, params) if mode != "ostromoukhov" else None; dt = time.perf_counter() - t0
            if ref is not None: print("   oracle (C port) %.1f ms, mismatches %d" % (dt * 1e3, int((out != ref).any(axis=2).sum())))
