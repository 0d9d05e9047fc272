"""Headline step (Floyd-Steinberg + Atkinson + JJN over the same batch) on ONE stream against one
stream per variant: does the next kernel's head fill the previous kernel's drain tail?

    [DP_WAVE_WARPS=8] python tools/wave_streams.py [--frames 128]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    from dither_pie_b200 import _capi, engine, synth
    _capi.ensure_device()
    L = _capi.lib()
    pal = engine.get_palette(synth.random_palette(256))
    h, w = 2160, 3840
    base = np.stack([synth.frame(h, w, 1 + t) for t in range(4)])
    frames = np.concatenate([base] * ((a.frames + 3) // 4))[:a.frames]
    src = _capi.DeviceBuffer(frames.nbytes).upload(np.ascontiguousarray(frames))
    variants = ("floyd_steinberg", "atkinson", "jjn")
    plans = [engine.Plan("error_diffusion", {"variant": v}, h, w) for v in variants]
    dsts = [_capi.DeviceBuffer(frames.nbytes) for _ in variants]
    streams = []
    for _ in variants:
        s = C.c_void_p()
        _capi.check(L.dp_stream_create(C.byref(s)), "dp_stream_create")
        streams.append(s)
    px = a.frames * h * w * len(variants)
    for label, per_variant in (("one stream", False), ("stream per variant", True), ("jjn first, stream per variant", True)):
        order = [2, 0, 1] if label.startswith("jjn") else [0, 1, 2]
        ts = []
        for r in range(a.reps):
            _capi.sync()
            for s in streams:
                _capi.check(L.dp_stream_sync(s), "sync")
            t0 = time.perf_counter()
            for i in order:
                plans[i].run(pal, src.ptr, a.frames, dsts[i].ptr, None, streams[i] if per_variant else streams[0])
            for s in streams:
                _capi.check(L.dp_stream_sync(s), "sync")
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[1:])
        dt = ts[len(ts) // 2]
        print(f"{label:32s} warps={os.environ.get('DP_WAVE_WARPS', 'default')}: {dt*1e3:.2f} ms per step "
              f"({px/dt/1e9:.2f} Gpx/s)")


if __name__ == "__main__":
    main()
