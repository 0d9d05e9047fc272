"""Small driver for ncu: runs one mode a few times on device-resident synthetic frames.

    python tools/prof_driver.py --mode bayer --h 1080 --w 1920 --frames 64 --k 16 --reps 3
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bayer")
    ap.add_argument("--params", default="{}")
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--palette", default="auto")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--pixelize", type=int, default=0)
    ap.add_argument("--upscale", type=int, default=1)
    a = ap.parse_args()
    import json
    from dither_pie_b200 import _capi, engine, synth
    _capi.ensure_device()
    params = json.loads(a.params)
    if a.palette == "pico8" or (a.palette == "auto" and a.k == 16):
        pal_rows = synth.hex_palette(synth.PICO8)
    elif a.palette == "c64":
        pal_rows = synth.hex_palette(synth.C64)
    else:
        pal_rows = synth.random_palette(a.k)
    pal = engine.get_palette(pal_rows)
    base = np.stack([synth.frame(a.h, a.w, t) for t in range(2)])
    frames = np.concatenate([base] * ((a.frames + 1) // 2))[:a.frames]
    src = _capi.DeviceBuffer(frames.nbytes).upload(np.ascontiguousarray(frames))
    if a.pixelize:
        tw, th = engine.even_dimensions(a.w, a.h, a.pixelize)
        plan = engine.Plan(a.mode, params, th, tw, (a.h, a.w), a.upscale)
    else:
        plan = engine.Plan(a.mode, params, a.h, a.w)
    dst = _capi.DeviceBuffer(a.frames * plan.out_h * plan.out_w * 3)
    import time
    ts = []
    for r in range(a.reps):
        _capi.sync()
        t0 = time.perf_counter()
        plan.run(pal, src.ptr, a.frames, dst.ptr, None, None)
        _capi.sync()
        ts.append(time.perf_counter() - t0)
    px = a.frames * a.h * a.w
    ts = sorted(ts[1:] or ts)
    dt = ts[len(ts) // 2]
    print(f"{a.mode} {a.params} {a.h}x{a.w} x{a.frames} K={a.k}: median {dt*1e3:.3f} ms (min {ts[0]*1e3:.3f}, max "
          f"{ts[-1]*1e3:.3f}, n={len(ts)})  {px/dt/1e9:.2f} Gpx/s  {6*px/dt/1e9:.1f} GB/s(alg)")


if __name__ == "__main__":
    main()
