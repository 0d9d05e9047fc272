"""Debug: per-phase cycle counts of the persistent k-means loop kernel (block 1, thread 0).
    python -m dither_pie_b200.build --timing
    gpurun -- 'python tools/km_timing.py'"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import _capi, kmeans, synth
_capi._lib = _capi.load_library(os.path.join(os.path.dirname(_capi.LIB_PATH), "_obj", "libditherpie_b200_timing.so"))
_capi._lib.dp_debug_km_timing.argtypes = [C.c_void_p, C.c_int]
_capi.ensure_device()
img = synth.frame(2160, 3840, 2).reshape(-1, 3)
init = img[np.random.RandomState(0).choice(len(img), 16, replace=False)].astype(np.float64)
buf = _capi.DeviceBuffer(img.nbytes).upload(np.ascontiguousarray(img))
out = (C.c_ulonglong * 8)()
names = ["totals+centres", "grid build", "barrier 1", "fill smem", "assign", "barrier 2", "loop top"]
for n in (len(img), len(img) // 8, 1024):
    kmeans.lloyd_device(buf.ptr, n, init, -1.0, 3)
    _capi.lib().dp_debug_km_timing(out, 1)
    iters = 50
    kmeans.lloyd_device(buf.ptr, n, init, -1.0, iters)
    _capi.lib().dp_debug_km_timing(out, 1)
    print(f"n={n}: cycles per iteration (block 1): " + ", ".join(f"{nm} {out[i]/iters:.0f}" for i, nm in enumerate(names)))
