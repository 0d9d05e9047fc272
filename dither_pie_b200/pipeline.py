"""Frame pipeline with HOST buffers: the GPU-side replacement of the reference's batch loop
(video_processor.py:304-346: ``Pool(num_workers).map`` over 15-frame batches of PNG files).

A clip is cut into batches of frames; batch n+1 is copied in and batch n-1 copied out while batch
n is dithered:

    copy-in stream   H2D of the batch (pinned host memory -> device)
    kernel stream    one launch per plan (several dither modes can share one upload)
    copy-out stream  D2H of every plan's result (colour bytes, or the 1-byte index plane)

with three device buffer sets and CUDA events between the streams.  Host arrays that are
page-locked (``pinned_empty`` / ``PinnedArray`` / ``dp_host_register``) are the DMA source and
target themselves; ordinary numpy arrays are staged through a pinned ring by the calling thread
(a host memcpy per batch, which then bounds the rate -- hand in pinned arrays for speed).

Everything goes through the C ABI (dp_stream_*, dp_event_*, dp_memcpy_*, the dither entry points).
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _capi, engine
from ._capi import DeviceBuffer, PinnedArray, check, lib


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array in page-locked host memory (kept alive by the array's ``base`` chain)."""
    pa = PinnedArray(shape, dtype)
    arr = pa.array
    _PINNED_OWNERS[arr.ctypes.data] = pa     # the view does not own the allocation
    return arr


_PINNED_OWNERS = {}


def release_pinned(arr: np.ndarray) -> None:
    pa = _PINNED_OWNERS.pop(arr.ctypes.data, None)
    if pa is not None:
        pa.free()


def is_pinned(arr: np.ndarray) -> bool:
    if arr.nbytes == 0:
        return False
    flag = C.c_int(0)
    check(lib().dp_host_is_pinned(arr.ctypes.data, C.byref(flag)), "dp_host_is_pinned")
    return bool(flag.value)


class _Event:
    def __init__(self, timing: bool = False):
        h = C.c_void_p()
        check(lib().dp_event_create(C.byref(h), int(timing)), "dp_event_create")
        self.h = h.value

    def record(self, stream):
        check(lib().dp_event_record(self.h, stream), "dp_event_record")

    def sync(self):
        check(lib().dp_event_sync(self.h), "dp_event_sync")

    def free(self):
        if self.h:
            lib().dp_event_destroy(self.h)
            self.h = None


class _Stream:
    def __init__(self):
        h = C.c_void_p()
        check(lib().dp_stream_create(C.byref(h)), "dp_stream_create")
        self.h = h.value

    def wait(self, ev: _Event):
        check(lib().dp_stream_wait_event(self.h, ev.h), "dp_stream_wait_event")

    def sync(self):
        check(lib().dp_stream_sync(self.h), "dp_stream_sync")

    def free(self):
        if self.h:
            lib().dp_stream_destroy(self.h)
            self.h = None


class FramePipeline:
    """``plans``: engine.Plan objects with the same source size (one per dither mode applied to
    every frame).  ``output``: "rgb" (colour bytes, [F, out_h, out_w, 3], what the reference's
    frames are), "index" (palette rows, u8 [F, h, w] -- 1 byte per dithered pixel) or "both".

    ``run(frames)`` processes a clip and returns; ``submit(frames, ...)`` only enqueues (the call
    returns while the GPU works, consecutive submits keep the three streams busy across calls) and
    ``flush()`` waits for everything submitted so far."""

    SLOTS = int(os.environ.get("DP_PIPE_SLOTS", "3"))    # device buffer sets in flight

    def __init__(self, plans: Sequence[engine.Plan], pal: engine.PaletteHandle, batch_frames: int,
                 output: str = "rgb"):
        assert output in ("rgb", "index", "both")
        _capi.ensure_device()
        self.plans = list(plans)
        self.pal = pal
        self.B = int(batch_frames)
        self.output = output
        p0 = self.plans[0]
        self.src_h, self.src_w = p0.src_h, p0.src_w
        for p in self.plans:
            assert (p.src_h, p.src_w) == (self.src_h, self.src_w), "plans must share the source size"
        self.in_frame = self.src_h * self.src_w * 3
        self.rgb_frame = [p.out_h * p.out_w * 3 for p in self.plans]
        self.idx_frame = [p.h * p.w for p in self.plans]
        self.want_rgb = output in ("rgb", "both")
        self.want_idx = output in ("index", "both")
        S, B, nv = self.SLOTS, self.B, len(self.plans)
        # one kernel stream per plan: when a launch drains (its last row bands), the next plan's
        # blocks take over the SMs that fall idle
        self.s_in, self.s_out = _Stream(), _Stream()
        self.s_k = [_Stream() for _ in range(nv)]
        self.d_src = [DeviceBuffer(max(B * self.in_frame, 16)) for _ in range(S)]
        self.d_rgb = [[DeviceBuffer(max(B * n, 16)) if self.want_rgb else None for n in self.rgb_frame]
                      for _ in range(S)]
        self.d_idx = [[DeviceBuffer(max(B * n, 16)) if self.want_idx else None for n in self.idx_frame]
                      for _ in range(S)]
        self.ev_in = [_Event() for _ in range(S)]
        self.ev_k = [[_Event() for _ in self.plans] for _ in range(S)]
        self.ev_out = [[_Event() for _ in self.plans] for _ in range(S)]
        self.stage_in: List[Optional[PinnedArray]] = [None] * S      # allocated on first use
        self.stage_rgb = [[None] * nv for _ in range(S)]
        self.stage_idx = [[None] * nv for _ in range(S)]
        self.n = 0                 # batches submitted so far (slot = n % SLOTS)
        self.pending = {}          # batch number -> [(stage array, destination view)]
        self.stats = {}
        self._reset_stats()

    def _reset_stats(self):
        self._t0 = None
        self._frames = self._h2d = self._d2h = 0
        self._pinned_in = self._pinned_out = True

    # ------------------------------------------------------------------------------------
    def close(self):
        for s in (self.s_in, self.s_out, *self.s_k):
            try:
                s.sync()
            except Exception:
                pass
        for group in (self.d_src, *self.d_rgb, *self.d_idx):
            for b in group:
                if b is not None:
                    b.free()
        for group in (self.stage_in, *self.stage_rgb, *self.stage_idx):
            for b in group:
                if b is not None:
                    b.free()
        for e in (*self.ev_in, *[e for r in self.ev_k for e in r], *[e for r in self.ev_out for e in r]):
            e.free()
        for s in (self.s_in, self.s_out, *self.s_k):
            s.free()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------------------------
    def _stage(self, table, slot, v, nbytes):
        if table[slot][v] is None:
            table[slot][v] = PinnedArray((nbytes,), np.uint8)
        return table[slot][v]

    def _drain(self, m):
        """wait for batch m's copies-out and move staged results to the caller's arrays"""
        jobs = self.pending.pop(m, None)
        if not jobs:              # results went to pinned arrays by DMA: nothing to do on the host
            return
        slot = m % self.SLOTS
        for v in range(len(self.plans)):
            self.ev_out[slot][v].sync()
        for stage, dest in jobs:
            np.copyto(dest, stage.array[:dest.nbytes].reshape(dest.shape))

    def submit(self, frames: np.ndarray, out_rgb: Optional[Sequence[np.ndarray]] = None,
               out_idx: Optional[Sequence[np.ndarray]] = None, progress=None):
        """Enqueue u8 [F, src_h, src_w, 3] host frames.  ``out_rgb`` / ``out_idx``: one destination
        array per plan (required for the planes the pipeline produces; pinned ones are written by
        DMA directly).  The arrays must stay alive and untouched until ``flush()`` returns."""
        L = lib()
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
        F = int(frames.shape[0])
        assert frames.shape[1:] == (self.src_h, self.src_w, 3), frames.shape
        nv = len(self.plans)
        assert not self.want_rgb or (out_rgb is not None and len(out_rgb) == nv)
        assert not self.want_idx or (out_idx is not None and len(out_idx) == nv)
        if F == 0:
            return
        in_pinned = is_pinned(frames)
        rgb_pinned = [self.want_rgb and is_pinned(out_rgb[v]) for v in range(nv)]
        idx_pinned = [self.want_idx and is_pinned(out_idx[v]) for v in range(nv)]
        self._pinned_in &= in_pinned
        self._pinned_out &= all(rgb_pinned[v] or not self.want_rgb for v in range(nv)) and \
            all(idx_pinned[v] or not self.want_idx for v in range(nv))
        if self._t0 is None:
            self._t0 = time.perf_counter()
        S, B = self.SLOTS, self.B
        for lo in range(0, F, B):
            hi = min(F, lo + B)
            cnt = hi - lo
            n = self.n
            slot = n % S
            if n >= S:
                self._drain(n - S)      # frees this slot's output staging
            # ---- copy in (the kernels of batch n - S have released the source buffer)
            if n >= S:
                for v in range(nv):
                    self.s_in.wait(self.ev_k[slot][v])
            nbytes = cnt * self.in_frame
            if in_pinned:
                src_host = frames[lo:hi].ctypes.data
            else:
                if self.stage_in[slot] is None:
                    self.stage_in[slot] = PinnedArray((B * self.in_frame,), np.uint8)
                elif n >= S:
                    self.ev_in[slot].sync()       # the H2D that last read this stage has finished
                st = self.stage_in[slot]
                np.copyto(st.array[:nbytes].reshape(frames[lo:hi].shape), frames[lo:hi])
                src_host = st.ptr
            check(L.dp_memcpy_h2d(self.d_src[slot].ptr, src_host, nbytes, self.s_in.h), "h2d")
            self._h2d += nbytes
            self.ev_in[slot].record(self.s_in.h)
            jobs = []
            for v, plan in enumerate(self.plans):
                sk = self.s_k[v]
                # ---- kernel
                sk.wait(self.ev_in[slot])
                if n >= S:
                    sk.wait(self.ev_out[slot][v])      # the previous result has left the buffers
                rgb = self.d_rgb[slot][v]
                idx = self.d_idx[slot][v]
                plan.run(self.pal, self.d_src[slot].ptr, cnt, rgb.ptr if rgb else None,
                         idx.ptr if idx else None, sk.h)
                self.ev_k[slot][v].record(sk.h)
                # ---- copy out
                self.s_out.wait(self.ev_k[slot][v])
                if self.want_rgb:
                    nby = cnt * self.rgb_frame[v]
                    dest = out_rgb[v][lo:hi]
                    if rgb_pinned[v]:
                        check(L.dp_memcpy_d2h(dest.ctypes.data, rgb.ptr, nby, self.s_out.h), "d2h")
                    else:
                        stg = self._stage(self.stage_rgb, slot, v, B * self.rgb_frame[v])
                        check(L.dp_memcpy_d2h(stg.ptr, rgb.ptr, nby, self.s_out.h), "d2h")
                        jobs.append((stg, dest))
                    self._d2h += nby
                if self.want_idx:
                    nby = cnt * self.idx_frame[v]
                    dest = out_idx[v][lo:hi]
                    if idx_pinned[v]:
                        check(L.dp_memcpy_d2h(dest.ctypes.data, idx.ptr, nby, self.s_out.h), "d2h")
                    else:
                        stg = self._stage(self.stage_idx, slot, v, B * self.idx_frame[v])
                        check(L.dp_memcpy_d2h(stg.ptr, idx.ptr, nby, self.s_out.h), "d2h")
                        jobs.append((stg, dest))
                    self._d2h += nby
                self.ev_out[slot][v].record(self.s_out.h)
            self.pending[n] = jobs
            self.n += 1
            self._frames += cnt
            if progress is not None:
                progress(hi, F)

    def flush(self):
        """Wait until every submitted batch has been delivered to the caller's arrays."""
        for m in sorted(self.pending):
            self._drain(m)
        for s in (self.s_in, self.s_out, *self.s_k):
            s.sync()
        dt = (time.perf_counter() - self._t0) if self._t0 is not None else 0.0
        self.stats = {"frames": self._frames, "seconds": dt, "h2d_bytes": self._h2d, "d2h_bytes": self._d2h,
                      "h2d_gbs": self._h2d / dt / 1e9 if dt > 0 else 0.0,
                      "d2h_gbs": self._d2h / dt / 1e9 if dt > 0 else 0.0,
                      "in_pinned": self._pinned_in, "out_pinned": self._pinned_out}
        self._reset_stats()

    def run(self, frames: np.ndarray, out_rgb: Optional[Sequence[np.ndarray]] = None,
            out_idx: Optional[Sequence[np.ndarray]] = None, progress=None):
        """frames: u8 [F, src_h, src_w, 3] on the host.  Returns (rgb_list, idx_list): one array
        per plan (None for a plane that was not asked for).  ``out_rgb`` / ``out_idx`` supply the
        destination arrays (pinned ones are written by DMA directly)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        F = int(frames.shape[0])
        nv = len(self.plans)
        if self.want_rgb and out_rgb is None:
            out_rgb = [np.empty((F, p.out_h, p.out_w, 3), np.uint8) for p in self.plans]
        if self.want_idx and out_idx is None:
            out_idx = [np.empty((F, p.h, p.w), np.uint8) for p in self.plans]
        self.submit(frames, out_rgb, out_idx, progress)
        self.flush()
        return (list(out_rgb) if self.want_rgb else [None] * nv,
                list(out_idx) if self.want_idx else [None] * nv)
