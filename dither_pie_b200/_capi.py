"""ctypes binding of libditherpie_b200.so (the C ABI declared in include/ditherpie_b200.h).

This is the stub a maintainer of the reference would add.  There is no
fallback: if the library is missing or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libditherpie_b200.so")

_lib = None
_lock = threading.Lock()


class DitherPieError(RuntimeError):
    pass


class Geometry(C.Structure):
    _fields_ = [("src_h", C.c_int32), ("src_w", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("upscale", C.c_int32), ("ytab", C.c_void_p), ("xtab", C.c_void_p)]


_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES)
PROTOTYPES = {
    "dp_last_error": [],
    "dp_version": [],
    "dp_device_count": [C.POINTER(C.c_int)],
    "dp_set_device": [_i],
    "dp_malloc": [C.POINTER(_vp), _sz],
    "dp_free": [_vp],
    "dp_host_alloc": [C.POINTER(_vp), _sz],
    "dp_host_free": [_vp],
    "dp_memcpy_h2d": [_vp, _vp, _sz, _vp],
    "dp_memcpy_d2h": [_vp, _vp, _sz, _vp],
    "dp_memset": [_vp, _i, _sz, _vp],
    "dp_stream_create": [C.POINTER(_vp)],
    "dp_stream_destroy": [_vp],
    "dp_stream_sync": [_vp],
    "dp_event_create": [C.POINTER(_vp), _i],
    "dp_event_destroy": [_vp],
    "dp_event_record": [_vp, _vp],
    "dp_stream_wait_event": [_vp, _vp],
    "dp_event_sync": [_vp],
    "dp_event_elapsed_ms": [_vp, _vp, C.POINTER(C.c_float)],
    "dp_host_register": [_vp, _sz],
    "dp_host_unregister": [_vp],
    "dp_host_is_pinned": [_vp, C.POINTER(C.c_int)],
    "dp_range_push": [C.c_char_p],
    "dp_range_pop": [],
    "dp_palette_create": [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                          C.POINTER(_vp)],
    "dp_palette_destroy": [_vp],
    "dp_palette_num_colors": [_vp],
    "dp_threshold_dither": [_vp, _vp, _i, C.POINTER(Geometry), _i, _vp, _i, _i, _f, _f, _f,
                            _vp, _vp, _vp],
    "dp_halftone": [_vp, _vp, _i, _i, _i, _i, _d, _d, _d, _d, _d, _i, _d, _vp, _vp, _vp, _vp],
    "dp_error_diffusion": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "dp_ostromoukhov": [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp],
    "dp_hybrid": [_vp, _vp, _i, _i, _i, _d, _d, _vp, _vp, _vp],
    "dp_perceptual": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "dp_adaptive_variance": [_vp, _vp, _i, _i, _i, _d, _i, _vp, _vp, _vp],
    "dp_unique_colors_pyset_order": [_vp, _i64, _vp, _vp],
    "dp_blue_noise_from_order": [_vp, _i, _vp],
    "dp_resample_nearest": [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp],
    "dp_kmeans_accumulate": [_vp, _i64, _vp, _i, _vp, _vp],
    "dp_kmeans_update": [_vp, _i, _vp, _vp, _vp],
    "dp_kmeans_lloyd": [_vp, _i64, _vp, _i, _d, _i, _vp, _i, C.POINTER(_i), C.POINTER(_d),
                        C.POINTER(C.c_ulonglong), C.POINTER(_i), _vp],
    "dp_kmeans_lloyd_p2p": [_vp, _i64, _vp, _i, _d, _i, _i, _i, _vp, C.c_ulonglong, _i, C.POINTER(_i),
                            C.POINTER(_d), C.POINTER(C.c_ulonglong), C.POINTER(_i), _vp],
    "dp_p2p_inbox_bytes": [],
    "dp_p2p_alloc": [_sz, C.POINTER(_vp), _vp],
    "dp_p2p_open": [_vp, C.POINTER(_vp)],
    "dp_p2p_close": [_vp],
    "dp_p2p_free": [_vp],
    "dp_nccl_load": [C.c_char_p],
    "dp_nccl_unique_id": [_vp],
    "dp_nccl_comm_create": [_vp, _i, _i, C.POINTER(_vp)],
    "dp_nccl_comm_destroy": [_vp],
    "dp_nccl_allreduce_u64": [_vp, _sz, _vp, _vp],
    "dp_threshold_dither_host": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _f, _f, _f, _vp],
}
_RESTYPES = {"dp_last_error": C.c_char_p}


def load_library(path: str = LIB_PATH):
    """dlopen the library and bind every declared symbol (no CUDA call is made)."""
    if not os.path.exists(path):
        raise DitherPieError(
            f"{path} is missing: build it with `python -m dither_pie_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    return lib


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                _lib = load_library()
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dp_last_error()
        raise DitherPieError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


_device_ready = threading.local()


def ensure_device(device: int | None = None):
    """Select the CUDA device for the calling thread (default: LOCAL_RANK or 0) and verify it
    is an sm_100a part.  Raises if there is no GPU -- the product path never runs on the CPU."""
    want = device
    if want is None:
        want = getattr(_device_ready, "dev", None)
        if want is not None:
            return want
        want = int(os.environ.get("LOCAL_RANK", "0"))
    n = C.c_int(0)
    check(lib().dp_device_count(C.byref(n)), "dp_device_count")
    if n.value < 1:
        raise DitherPieError("no CUDA device visible; dither_pie_b200 has no CPU fallback")
    want %= n.value
    check(lib().dp_set_device(want), "dp_set_device")
    _device_ready.dev = want
    return want


class DeviceBuffer:
    """Owned device allocation (cudaMalloc) with explicit free."""

    def __init__(self, nbytes: int):
        ensure_device()
        p = C.c_void_p()
        check(lib().dp_malloc(C.byref(p), nbytes), "dp_malloc")
        self.ptr = p.value
        self.nbytes = nbytes

    def upload(self, arr, stream=None):
        assert arr.flags["C_CONTIGUOUS"] and arr.nbytes <= self.nbytes
        check(lib().dp_memcpy_h2d(self.ptr, arr.ctypes.data, arr.nbytes, stream), "h2d")
        return self

    def download(self, arr, stream=None):
        assert arr.flags["C_CONTIGUOUS"] and arr.nbytes <= self.nbytes
        check(lib().dp_memcpy_d2h(arr.ctypes.data, self.ptr, arr.nbytes, stream), "d2h")
        return arr

    def free(self):
        if self.ptr:
            lib().dp_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sync(stream=None):
    check(lib().dp_stream_sync(stream), "dp_stream_sync")


class PinnedArray:
    """numpy view of pinned (page-locked) host memory; free() or garbage collection releases it."""

    def __init__(self, shape, dtype):
        import numpy as np

        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ensure_device()
        p = C.c_void_p()
        check(lib().dp_host_alloc(C.byref(p), max(n, 1)), "dp_host_alloc")
        self.ptr = p.value
        self._buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(self._buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._buf = None
            lib().dp_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
