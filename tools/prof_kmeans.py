"""Small driver for ncu: k-means assignment passes over device-resident synthetic pixels.

    python tools/prof_kmeans.py --frames 16 --k 16 --reps 3
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16, help="distinct 4K frames in the pixel array")
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--loop", type=int, default=0, help="instead: the persistent Lloyd loop, this many iterations "
                                                        "over ALL the pixels")
    a = ap.parse_args()
    from dither_pie_b200 import _capi, synth
    from dither_pie_b200._capi import DeviceBuffer, check, lib
    _capi.ensure_device()
    px = np.concatenate([synth.frame(2160, 3840, 2 + t).reshape(-1, 3) for t in range(a.frames)])
    n = px.shape[0]
    rs = np.random.RandomState(0)
    init = px[rs.choice(n, a.k, replace=False)].astype(np.float64)
    buf = DeviceBuffer(px.nbytes).upload(np.ascontiguousarray(px))
    if a.loop:
        from dither_pie_b200 import kmeans
        kmeans.lloyd_device(buf.ptr, n, init, -1.0, 2)
        ts = []
        for _ in range(a.reps):
            _capi.sync()
            t0 = time.perf_counter()
            kmeans.lloyd_device(buf.ptr, n, init, -1.0, a.loop)
            ts.append(time.perf_counter() - t0)
        dt = sorted(ts)[len(ts) // 2]
        print(f"kmeans Lloyd loop K={a.k} {n/1e6:.1f} Mpx x {a.loop} iterations: median {dt*1e3:.3f} ms  "
              f"{dt/a.loop*1e6:.1f} us/iteration  {3*n*a.loop/dt/1e9:.0f} GB/s(alg)")
        return
    cent = DeviceBuffer(init.nbytes).upload(init)
    sums = DeviceBuffer((a.k * 4 + 1) * 8)
    ts = []
    for _ in range(a.reps):
        check(lib().dp_memset(sums.ptr, 0, (a.k * 4 + 1) * 8, None))
        _capi.sync()
        t0 = time.perf_counter()
        check(lib().dp_kmeans_accumulate(buf.ptr, n, cent.ptr, a.k, sums.ptr, None))
        _capi.sync()
        ts.append(time.perf_counter() - t0)
    dt = sorted(ts)[len(ts) // 2]
    print(f"kmeans assign K={a.k} {n/1e6:.1f} Mpx: median {dt*1e3:.3f} ms  {n/dt/1e9:.1f} Gpx/s  {3*n/dt/1e9:.0f} GB/s(alg)")


if __name__ == "__main__":
    main()
