"""Debug: per-phase cycle counts of k_diffuse_wave.  Needs the instrumented library:
    python -m dither_pie_b200.build --timing
    gpurun -- 'python tools/wave_timing.py floyd_steinberg 1'"""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import _capi, engine, synth
_capi._lib = _capi.load_library(os.path.join(os.path.dirname(_capi.LIB_PATH), "_obj", "libditherpie_b200_timing.so"))
_capi._lib.dp_debug_wave_timing.argtypes = [C.c_void_p, C.c_int]
_capi.ensure_device()
L = _capi.lib()
variant = sys.argv[1] if len(sys.argv) > 1 else "floyd_steinberg"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
h, w = 2160, 3840
pal = engine.get_palette(synth.random_palette(K))
img = np.stack([synth.frame(h, w, 1 + t % 2) for t in range(frames)])
src = _capi.DeviceBuffer(img.nbytes).upload(img)
dst = _capi.DeviceBuffer(img.nbytes)
if variant in ("perceptual", "ostromoukhov", "hybrid", "adaptive_variance"):
    plan = engine.Plan(variant, {}, h, w)      # the other wavefront instantiations
else:
    plan = engine.Plan("error_diffusion", {"variant": variant}, h, w)
buf = (C.c_ulonglong * 532)()
for rep in range(2):
    plan.run(pal, src.ptr, frames, dst.ptr, None, None)
    _capi.sync()
    L.dp_debug_wave_timing(buf, 1)
cnt = buf[512], buf[513]
t = np.array(buf[:512], dtype=np.float64).reshape(128, 4)
nch = (w + 64 + 31) // 32
print(variant, "frames", frames, "K", K, " cycles per chunk (wait, stage, steps, writeback):")
for b in (0, 1, 2, 10, 30, 60, 67):
    print(b, (t[b] / nch).round(0))
print("mean over bands 1..66:", (t[1:67].mean(0) / nch).round(0), " per step:", round(t[1:67, 2].mean() / nch / 32, 1))
print("slow-path pixels:", cnt[0], " warp-steps with a slow lane:", cnt[1], " of", frames * 68 * nch * 32, "warp-steps")
st = np.array(buf[516:532], dtype=np.float64)
print("band 10 lane 5 per-step cycles [recv+feed, acc+clamp, search, e+taps, out, emit/shift]:", (st[:6] / (nch * 32)).round(1))
