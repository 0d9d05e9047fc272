cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import numpy as np, time, ctypes as C
from dither_pie_b200 import _capi, synth
from dither_pie_b200._capi import check, lib, DeviceBuffer
_capi.ensure_device()
img = synth.frame(2160, 3840, 2).reshape(-1, 3)
n = len(img)
buf = DeviceBuffer(img.nbytes).upload(np.ascontiguousarray(img))
for K in (16, 8, 64):
    cent = img[:: n // K][:K].astype(np.float64)
    cdev = DeviceBuffer(K * 3 * 8).upload(cent.copy())
    sums = DeviceBuffer(K * 4 * 8)
    ts = []
    for r in range(8):
        check(lib().dp_memset(sums.ptr, 0, K * 4 * 8, None)); _capi.sync()
        t0 = time.perf_counter()
        check(lib().dp_kmeans_accumulate(buf.ptr, n, cdev.ptr, K, sums.ptr, None)); _capi.sync()
        ts.append(time.perf_counter() - t0)
    dt = sorted(ts[1:])[3]
    print(f"kmeans accumulate K={K}: {dt*1e3:.3f} ms  {n/dt/1e9:.1f} Gpx/s  {3*n/dt/1e9:.0f} GB/s")
PY
