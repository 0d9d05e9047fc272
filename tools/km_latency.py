"""Lloyd loop latency: marginal cost per iteration (100 vs 20 iterations) for several pixel counts.
    gpurun -- python tools/km_latency.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import _capi, kmeans, synth
_capi.ensure_device()
img = synth.frame(2160, 3840, 2).reshape(-1, 3)
init = img[np.random.RandomState(0).choice(len(img), 16, replace=False)].astype(np.float64)
buf = _capi.DeviceBuffer(img.nbytes).upload(np.ascontiguousarray(img))
for n in (len(img), len(img) // 2, len(img) // 8, 65536, 1024):
    for K, ini in ((16, init), (40, np.concatenate([init, init[:8] + 1.5, init + 3.25])[:40])):
        ts = {}
        for iters in (20, 100):
            kmeans.lloyd_device(buf.ptr, n, ini, -1.0, 2)
            best = 1e9
            for _ in range(3):
                _capi.sync()
                t0 = time.perf_counter()
                kmeans.lloyd_device(buf.ptr, n, ini, -1.0, iters, check_every=iters)
                best = min(best, time.perf_counter() - t0)
            ts[iters] = best
        print(f"n={n:8d} K={K}: 20 it {ts[20]*1e3:.3f} ms, 100 it {ts[100]*1e3:.3f} ms, marginal {(ts[100]-ts[20])/80*1e6:.1f} us/it, fixed {(ts[20]-(ts[100]-ts[20])/4)*1e3:.3f} ms")
