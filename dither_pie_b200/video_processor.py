"""Drop-in surface of the reference's ``video_processor`` for the regular (non-neural) path.

pixelize_regular (video_processor.py:563-577), _apply_final_resize_to_frame (:393-420) and the
frame data-parallelism of VideoProcessor.process_video_streaming (:304-346) -- the reference's
``multiprocessing.Pool.map`` over 15-frame batches becomes contiguous frame shards over the
GPUs of one box (one process per GPU), each shard processed in large device batches by the
fused pixelize -> dither -> up-scale kernel.  FFmpeg extraction / re-encode (:98-170, :208-217,
:361-382) stays a subprocess outside the timed path.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path
from typing import Callable, Optional, Tuple

import numpy as np

from . import _capi, engine
from .dithering_lib import ImageDitherer, PixelizeMethod


def _compute_even_dimensions(orig_w: int, orig_h: int, max_size: int) -> Tuple[int, int]:
    """NeuralPixelizer._compute_even_dimensions (video_processor.py:547-560)."""
    return engine.even_dimensions(orig_w, orig_h, max_size)


def _resample_array(arr: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    arr = np.ascontiguousarray(arr, np.uint8)
    single = arr.ndim == 3
    if single:
        arr = arr[None]
    F, H, W, _ = arr.shape
    src = _capi.DeviceBuffer(max(arr.nbytes, 4)).upload(arr)
    dst = _capi.DeviceBuffer(max(F * out_h * out_w * 3, 4))
    try:
        engine.resample(src.ptr, F, H, W, out_h, out_w, dst.ptr)
        out = np.empty((F, out_h, out_w, 3), np.uint8)
        dst.download(out)
        _capi.sync()
    finally:
        src.free()
        dst.free()
    return out[0] if single else out


def pixelize_regular_array(arr: np.ndarray, max_size: int) -> np.ndarray:
    h, w = arr.shape[-3], arr.shape[-2]
    tw, th = engine.even_dimensions(w, h, max_size)
    return _resample_array(arr, th, tw)


def pixelize_regular(image, max_size: int):
    """PIL.Image -> PIL.Image with even dimensions (video_processor.py:563-577)."""
    from PIL import Image
    arr = np.array(image.convert('RGB'), dtype=np.uint8)
    return Image.fromarray(pixelize_regular_array(arr, max_size), 'RGB')


def _apply_final_resize_to_frame(image, multiplier: int):
    """Integer NEAREST up-scale, dimensions bumped to even (video_processor.py:393-420)."""
    from PIL import Image
    arr = np.array(image.convert('RGB'), dtype=np.uint8)
    h, w, _ = arr.shape
    nw, nh = w * multiplier, h * multiplier
    nw += nw % 2
    nh += nh % 2
    return Image.fromarray(_resample_array(arr, nh, nw), 'RGB')


def shard_frames(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range of ``rank`` (keeps output order trivial; no collective)."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class VideoProcessor:
    """video_processor.py:27-390.  ``num_workers`` is reinterpreted as the number of GPUs
    (= processes of the torchrun job); a single process drives its own shard."""

    def __init__(self, num_workers: Optional[int] = None,
                 progress_callback: Optional[Callable[[float, str], None]] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.num_workers = num_workers if num_workers is not None else self.world
        self.progress_callback = progress_callback

    def _report_progress(self, fraction: float, message: str):
        if self.progress_callback:
            self.progress_callback(fraction, message)

    # ---- array-level frame path (the timed path) ---------------------------------------
    def process_frames(self, frames: np.ndarray, ditherer: ImageDitherer,
                       pixelize_func=None, batch_size: int = 64,
                       final_resize_multiplier: Optional[int] = None) -> np.ndarray:
        """uint8 [F,H,W,3] -> uint8 [F,H',W',3]: the body of _process_single_frame
        (video_processor.py:443-462) for every frame of this process's shard, minus PNG I/O.
        ``pixelize_func`` is ``(method, max_size)`` or None as in the reference (:178)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        max_size = None
        if pixelize_func:
            method, max_size = pixelize_func
            method = getattr(method, "value", method)
            if method == PixelizeMethod.NEURAL.value:
                raise NotImplementedError("neural pixelization is outside the B200 hot path")
            if method != PixelizeMethod.REGULAR.value:
                max_size = None
        if ditherer.palette is None:
            first = frames[0]
            if max_size:
                first = pixelize_regular_array(first, max_size)
            ditherer._ensure_palette(first)
        if not ditherer.dither_mode:
            from .dithering_lib import DitherMode
            ditherer.dither_mode = DitherMode.NONE
        strategy = ditherer._get_dither_strategy(ditherer.dither_mode)
        outs = []
        n = frames.shape[0]
        for s in range(0, n, batch_size):
            outs.append(engine.dither_frames(
                frames[s:s + batch_size], ditherer.palette, strategy._mode,
                strategy.get_current_parameters(), use_gamma=ditherer.use_gamma,
                pixelize_max_size=max_size, final_multiplier=final_resize_multiplier,
                even_final=True))
            self._report_progress(min(1.0, (s + batch_size) / max(n, 1)),
                                  f"Processed {min(s + batch_size, n)}/{n} frames")
        return np.concatenate(outs, axis=0) if outs else frames[:0]

    # ---- file-level entry point (FFmpeg outside the timed path) -------------------------
    def process_video_streaming(self, input_path: str, output_path: str, ditherer: ImageDitherer,
                                pixelize_func=None, batch_size: int = 15,
                                final_resize_multiplier: Optional[int] = None) -> bool:
        """Same signature and bool contract as the reference (:172-178, :386-390)."""
        try:
            from PIL import Image
            if not shutil.which("ffmpeg") or not shutil.which("ffprobe"):
                raise RuntimeError("ffmpeg/ffprobe not found (video I/O is a subprocess)")
            with tempfile.TemporaryDirectory() as tmp:
                fps = subprocess.check_output(
                    ["ffprobe", "-v", "0", "-of", "csv=p=0", "-select_streams", "v:0",
                     "-show_entries", "stream=r_frame_rate", input_path]).decode().strip()
                subprocess.run(["ffmpeg", "-i", input_path, os.path.join(tmp, "frame_%05d.png")],
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                files = sorted(Path(tmp).glob("frame_*.png"))
                lo, hi = 0, len(files)  # file-level driver is single-process; shards use process_frames
                for s in range(lo, hi, max(batch_size, 1)):
                    chunk = files[s:min(s + batch_size, hi)]
                    arr = np.stack([np.array(Image.open(f).convert('RGB')) for f in chunk])
                    out = self.process_frames(arr, ditherer, pixelize_func, len(chunk),
                                              final_resize_multiplier)
                    for f, o in zip(chunk, out):
                        Image.fromarray(o, 'RGB').save(f)
                    self._report_progress((s - lo) / max(hi - lo, 1), "Processing frames")
                if True:
                    subprocess.run(["ffmpeg", "-y", "-framerate", fps, "-i",
                                    os.path.join(tmp, "frame_%05d.png"), "-i", input_path,
                                    "-map", "0:v", "-map", "1:a?", "-c:a", "copy", "-c:v",
                                    "libx264", "-crf", "18", "-pix_fmt", "yuv420p", output_path],
                                   check=True, stdout=subprocess.DEVNULL,
                                   stderr=subprocess.DEVNULL)
            self._report_progress(1.0, "Video processing complete!")
            return True
        except Exception as e:  # the reference's contract: report and return False
            self._report_progress(1.0, f"Error: {e}")
            print(f"Video processing error: {e}", file=sys.stderr)
            return False
