"""Drop-in surface of the reference's ``video_processor`` for the regular (non-neural) path.

pixelize_regular (video_processor.py:563-577), _apply_final_resize_to_frame (:393-420) and the
frame data-parallelism of VideoProcessor.process_video_streaming (:304-346).  The reference maps
15-frame batches of PNG files over a ``multiprocessing.Pool``; here

  * a process drives its GPU through ``pipeline.FramePipeline`` (pinned host memory, copy-in /
    kernel / copy-out streams, double-buffered device batches; fused pixelize -> dither -> up-scale
    kernel where the mode has one);
  * frames are sharded contiguously over ``num_workers`` GPUs: over the ranks of a torchrun job
    (RANK / WORLD_SIZE, one process per GPU, no data-path collective) or, in a single process,
    over one host thread per visible GPU;
  * the per-frame failure contract is kept: a frame is retried twice, then replaced by the nearest
    successfully processed frame (:325-336, :53-96);
  * frames travel between ffmpeg and the GPU as raw RGB through pipes instead of PNG files on disk
    (:208-217, :361-382) -- FFmpeg itself stays a subprocess outside the timed path.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import threading
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _capi, engine, pipeline
from .dithering_lib import DitherMode, ImageDitherer, PixelizeMethod


def _compute_even_dimensions(orig_w: int, orig_h: int, max_size: int) -> Tuple[int, int]:
    """NeuralPixelizer._compute_even_dimensions (video_processor.py:547-560)."""
    return engine.even_dimensions(orig_w, orig_h, max_size)


class NeuralPixelizer:
    """video_processor.py:478-560.  The neural pixelizer (a GAN whose weights the reference does
    not ship) is outside the B200 hot path; the class exists so that the reference's front ends,
    which import it unconditionally (dither_cli.py:26, dither_pie_gui.py:31), load against this
    module.  Constructing it is free; ``pixelize`` refuses loudly."""

    def __init__(self, device: Optional[str] = None):
        self._device = device

    def pixelize(self, image, max_size: int):
        raise NotImplementedError(
            "neural pixelization is outside the B200 hot path (SURVEY.md section 2 row 19); "
            "use the 'regular' pixelization method")

    @staticmethod
    def _compute_even_dimensions(orig_w: int, orig_h: int, max_size: int) -> Tuple[int, int]:
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


def _unpack_pixelize(pixelize_func):
    """``pixelize_func`` is ``(method, max_size)`` or None, as in the reference (:226-230)."""
    if not pixelize_func:
        return None
    method, max_size = pixelize_func
    method = getattr(method, "value", method)
    if method == PixelizeMethod.NEURAL.value:
        raise NotImplementedError("neural pixelization is outside the B200 hot path")
    return int(max_size) if method == PixelizeMethod.REGULAR.value else None


class VideoProcessor:
    """video_processor.py:27-390.  ``num_workers`` is the number of GPUs the frames are sharded
    over.  Default: every rank of the torchrun job (one process per GPU), or -- in a plain
    single process -- one GPU; ``num_workers=N`` in a single process drives N visible GPUs from N
    host threads."""

    DEVICE_BATCH = 32     # frames per device batch inside the pipeline

    def __init__(self, num_workers: Optional[int] = None,
                 progress_callback: Optional[Callable[[float, str], None]] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if num_workers is None:
            num_workers = self.world
        self.num_workers = max(1, int(num_workers))
        self.progress_callback = progress_callback
        self._pipes: dict = {}                   # (device, plan key) -> (plan, FramePipeline), reused
        self._pipes_lock = threading.Lock()
        self.failed_frames: List[int] = []       # indices (in the caller's array) fixed by copying
        self.last_range: Tuple[int, int] = (0, 0)  # frame range this process handled
        self.last_stats: dict = {}

    @staticmethod
    def _close_entry(entry):
        plan, pipe = entry[0], entry[1]
        try:
            pipe.close()
        finally:
            if hasattr(plan, "close"):
                plan.close()

    def close(self):
        """Release the cached pipelines (device buffers, streams)."""
        with self._pipes_lock:
            entries, self._pipes = list(self._pipes.values()), {}
        for e in entries:
            self._close_entry(e)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _report_progress(self, fraction: float, message: str):
        if self.progress_callback:
            self.progress_callback(fraction, message)

    # ---- one frame on the simple synchronous path: the unit of the retry contract ---------
    def _process_single_frame(self, frame: np.ndarray, ditherer: ImageDitherer, mode: str,
                              params: dict, max_size, final_resize_multiplier, output: str):
        """Body of the reference's _process_single_frame (video_processor.py:443-462) for one
        in-memory frame."""
        res = engine.dither_frames(frame, ditherer.palette, mode, params,
                                   use_gamma=ditherer.use_gamma, pixelize_max_size=max_size,
                                   final_multiplier=final_resize_multiplier, even_final=True,
                                   return_indices=(output == "index"))
        return res[1] if output == "index" else res

    @staticmethod
    def _fix_failed_frames(failed: List[int], out: np.ndarray) -> List[int]:
        """A failed frame becomes a copy of the nearest good one: previous frames first, then
        following ones (video_processor.py:53-96).  Returns the frames that could not be fixed."""
        bad = set(failed)
        lost = []
        for i in failed:
            src = next((j for j in range(i - 1, -1, -1) if j not in bad), None)
            if src is None:
                src = next((j for j in range(i + 1, out.shape[0]) if j not in bad), None)
            if src is None:
                print(f"ERROR: Could not find any successful frame to copy for frame {i}",
                      file=sys.stderr)
                lost.append(i)
                continue
            out[i] = out[src]
            print(f"Fixed frame {i} by copying from frame {src}", file=sys.stderr)
        return lost

    def _setup(self, frames, ditherer, pixelize_func):
        max_size = _unpack_pixelize(pixelize_func)
        if ditherer.palette is None:
            # one palette for the clip, fixed from its (pixelized) first frame; every rank derives
            # the same one from the same frame.  (The reference's CLI passes in a palette computed
            # from the FULL first frame, dither_cli.py:619-654; its workers, given None, would each
            # derive their own per frame.)
            first = frames[0]
            if max_size:
                first = pixelize_regular_array(first, max_size)
            ditherer._ensure_palette(first)
        if not ditherer.dither_mode:
            ditherer.dither_mode = DitherMode.NONE
        strategy = ditherer._get_dither_strategy(ditherer.dither_mode)
        return max_size, strategy._mode, strategy.get_current_parameters()

    def _run_range(self, frames, out, lo, hi, ditherer, mode, params, max_size, mult, output,
                   batch, report):
        """Frames [lo, hi) of ``frames`` -> ``out[lo:hi]`` on the calling thread's GPU."""
        if hi <= lo:
            return []
        H, W = frames.shape[1:3]
        pal = engine.get_palette(ditherer.palette, ditherer.use_gamma)
        key = (_capi.ensure_device(), mode, repr(sorted(params.items())), H, W, max_size, mult, batch,
               output, id(pal))
        try:
            # plan + pipeline (device buffers, streams, events) are kept between calls: a streamed
            # clip comes back chunk after chunk with the same geometry
            with self._pipes_lock:
                entry = self._pipes.pop(key, None)
            if entry is None:
                plan = engine.make_plan(mode, params, H, W, max_size, mult, True, max_frames=batch)
                entry = (plan, pipeline.FramePipeline([plan], pal, batch, output), pal)
            pipe = entry[1]
            try:
                kw = {"out_rgb": [out[lo:hi]]} if output == "rgb" else {"out_idx": [out[lo:hi]]}
                pipe.run(frames[lo:hi], progress=report, **kw)
                self.last_stats = dict(pipe.stats)
            except BaseException:
                self._close_entry(entry)
                raise
            with self._pipes_lock:
                self._pipes[key] = entry
                while len(self._pipes) > 4:
                    self._close_entry(self._pipes.pop(next(iter(self._pipes))))
            return []
        except NotImplementedError:
            raise
        except Exception as e:   # the reference's contract: retry per frame, then patch
            print(f"Batch path failed ({e}); retrying frame by frame...", file=sys.stderr)
        failed = []
        for i in range(lo, hi):
            ok = False
            for _ in range(3):          # first try + "retry up to 2 more times" (:331-335)
                try:
                    out[i] = self._process_single_frame(frames[i], ditherer, mode, params, max_size,
                                                        mult, output)
                    ok = True
                    break
                except Exception as e:
                    print(f"Error processing frame {i}: {e}", file=sys.stderr)
            if not ok:
                failed.append(i)
        return failed

    # ---- array-level frame path (the timed path) ---------------------------------------
    def process_frames(self, frames: np.ndarray, ditherer: ImageDitherer,
                       pixelize_func=None, batch_size: int = 64,
                       final_resize_multiplier: Optional[int] = None, output: str = "rgb",
                       out: Optional[np.ndarray] = None, shard: bool = True) -> np.ndarray:
        """uint8 [F,H,W,3] -> uint8 [F,H',W',3]: the body of _process_single_frame
        (video_processor.py:443-462) for every frame, minus PNG I/O.

        ``output="index"`` returns the palette-index plane u8 [F,h,w] instead of colour bytes.
        ``out``: destination array (page-locked memory from ``pipeline.pinned_empty`` is written
        by DMA directly).  In a torchrun job (WORLD_SIZE > 1) and with ``shard`` every rank
        processes only its contiguous range ``self.last_range`` and returns that part (use
        ``distributed.process_frames_sharded`` to gather on rank 0)."""
        assert output in ("rgb", "index")
        frames = frames if (isinstance(frames, np.ndarray) and frames.dtype == np.uint8
                            and frames.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(frames, np.uint8)
        n = int(frames.shape[0])
        self.failed_frames = []
        if n == 0:
            self.last_range = (0, 0)
            return frames[:0]
        max_size, mode, params = self._setup(frames, ditherer, pixelize_func)
        H, W = frames.shape[1:3]
        (h, w), (oh, ow), _ = engine.video_geometry(H, W, max_size, final_resize_multiplier, True)
        # device batch: about 256 MB of input per batch, never more than ``batch_size`` frames
        batch = max(1, min(int(batch_size), max(self.DEVICE_BATCH, (256 << 20) // (H * W * 3))))
        lo, hi = 0, n
        if shard and self.world > 1:
            workers = min(self.num_workers, self.world)
            lo, hi = shard_frames(n, self.rank, workers) if self.rank < workers else (0, 0)
        self.last_range = (lo, hi)
        shape = (hi - lo, oh, ow, 3) if output == "rgb" else (hi - lo, h, w)
        if out is None:
            out = np.empty(shape, np.uint8)
        assert out.shape == shape and out.dtype == np.uint8, (out.shape, shape)
        part = frames[lo:hi]
        m = hi - lo
        done = [0]
        lock = threading.Lock()

        def report(k, total):
            with lock:
                done[0] = max(done[0], k)
                self._report_progress(min(1.0, done[0] / max(m, 1)), f"Processed {done[0]}/{m} frames")

        threads_n = 1
        if self.world == 1 and self.num_workers > 1:
            cnt = _capi.C.c_int(0)
            _capi.check(_capi.lib().dp_device_count(_capi.C.byref(cnt)), "dp_device_count")
            threads_n = max(1, min(self.num_workers, cnt.value, m))
        failed: List[int] = []
        if threads_n == 1:
            failed = self._run_range(part, out, 0, m, ditherer, mode, params, max_size,
                                     final_resize_multiplier, output, batch, report)
        else:
            # one host thread per GPU; ctypes releases the GIL around every library call
            results = [None] * threads_n
            errors = []

            def work(t):
                try:
                    _capi.ensure_device(t)
                    a, b = shard_frames(m, t, threads_n)
                    results[t] = self._run_range(part, out, a, b, ditherer, mode, params, max_size,
                                                 final_resize_multiplier, output, batch, None)
                except Exception as e:      # pragma: no cover
                    errors.append(e)

            ts = [threading.Thread(target=work, args=(t,)) for t in range(threads_n)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errors:
                raise errors[0]
            failed = sorted(i for r in results if r for i in r)
            report(m, m)
        if failed:
            print(f"Fixing {len(failed)} failed frames by copying from nearest frames...",
                  file=sys.stderr)
            self._fix_failed_frames(failed, out)
            self.failed_frames = [lo + i for i in failed]
        return out

    # ---- file-level entry point (FFmpeg outside the timed path) -------------------------
    def get_video_info(self, video_path: str) -> dict:
        """fps / width / height / frame_count through ffprobe (video_processor.py:98-170)."""
        def probe(entries):
            r = subprocess.run(["ffprobe", "-v", "error", "-select_streams", "v:0", "-show_entries",
                                f"stream={entries}", "-of", "default=nokey=1:noprint_wrappers=1",
                                video_path], capture_output=True, text=True, check=True)
            return r.stdout.strip().split("\n")
        info = {'fps': 30.0, 'width': 1920, 'height': 1080, 'duration': None, 'frame_count': None}
        try:
            f = probe("r_frame_rate")[0]
            if "/" in f:
                a, b = f.split("/")
                info['fps'] = float(a) / float(b)
            elif f:
                info['fps'] = float(f)
            dims = probe("width,height")
            info['width'], info['height'] = int(dims[0]), int(dims[1])
            for line in probe("duration,nb_frames"):
                if line and line != 'N/A':
                    try:
                        v = float(line)
                    except ValueError:
                        continue
                    if v > 100:
                        info['frame_count'] = int(v)
                    else:
                        info['duration'] = v
            if info['frame_count'] is None and info['duration'] is not None:
                info['frame_count'] = int(info['duration'] * info['fps'])
        except Exception as e:
            print(f"Warning: Could not get video info: {e}", file=sys.stderr)
        return info

    def process_video_streaming(self, input_path: str, output_path: str, ditherer: ImageDitherer,
                                pixelize_func=None, batch_size: int = 15,
                                final_resize_multiplier: Optional[int] = None) -> bool:
        """Same signature and bool contract as the reference (:172-178, :386-390).

        Frames are read from ``ffmpeg -f rawvideo -pix_fmt rgb24`` through a pipe into pinned
        memory, dithered in chunks by the frame pipeline and written to the encoder's stdin -- no
        PNG files.  In a torchrun job every rank decodes and dithers its own contiguous frame
        range into ``<output>.parts/part_<rank>.rgb``; after a barrier rank 0 feeds the parts to
        the encoder in order.  ``palette=None``: one palette for the clip, from frame 0 after
        pixelization, identically on every rank (the reference's CLI passes a palette it computed
        from the full first frame, dither_cli.py:619-654; the reference's workers, given None, each
        derive their own per frame -- see INTEGRATION.md)."""
        try:
            if not shutil.which("ffmpeg") or not shutil.which("ffprobe"):
                raise RuntimeError("ffmpeg/ffprobe not found (video I/O is a subprocess)")
            info = self.get_video_info(input_path)
            fps, W, H = info['fps'], info['width'], info['height']
            total = info.get('frame_count') or 0
            self._report_progress(0.0, "Initializing video processing...")
            max_size = _unpack_pixelize(pixelize_func)
            (h, w), (oh, ow), _ = engine.video_geometry(H, W, max_size, final_resize_multiplier, True)
            workers = min(self.num_workers, self.world)
            if self.world > 1:
                if total <= 0:
                    raise RuntimeError("frame count unknown: cannot shard the clip over ranks")
                lo, hi = shard_frames(total, self.rank, workers) if self.rank < workers else (0, 0)
            else:
                lo, hi = 0, None
            if ditherer.palette is None:
                first = _read_raw_frames(input_path, W, H, 0, 1)
                if first.shape[0] == 0:
                    raise ValueError("No frames extracted from video")
                self._setup(first, ditherer, pixelize_func)
            chunk = max(int(batch_size), 4 * self.DEVICE_BATCH)
            parts_dir = output_path + ".parts"
            sink = None
            enc = None
            if self.world > 1:
                os.makedirs(parts_dir, exist_ok=True)
                sink = open(os.path.join(parts_dir, f"part_{self.rank:03d}.rgb"), "wb")
            else:
                enc = _open_encoder(output_path, input_path, fps, ow, oh)
                sink = enc.stdin
            count = 0
            try:
                if hi is None or hi > lo:
                    for arr in _iter_raw_frames(input_path, W, H, lo, hi, chunk):
                        out = self.process_frames(arr, ditherer, pixelize_func, self.DEVICE_BATCH,
                                                  final_resize_multiplier, shard=False)
                        sink.write(memoryview(out).cast("B"))
                        count += arr.shape[0]
                        denom = (hi - lo) if hi is not None else max(total, count)
                        self._report_progress(0.1 + 0.8 * min(1.0, count / max(denom, 1)),
                                              f"Processed {count}/{denom} frames")
            finally:
                sink.close()
            if self.world == 1 and count == 0:
                raise ValueError("No frames extracted from video")
            if self.world > 1:
                from . import distributed
                distributed.init_process_group()
                distributed.barrier()
                if self.rank == 0:
                    self._report_progress(0.9, "Encoding final video...")
                    enc = _open_encoder(output_path, input_path, fps, ow, oh)
                    for r in range(workers):
                        with open(os.path.join(parts_dir, f"part_{r:03d}.rgb"), "rb") as f:
                            shutil.copyfileobj(f, enc.stdin, 1 << 24)
                    enc.stdin.close()
                    rc = enc.wait()
                    shutil.rmtree(parts_dir, ignore_errors=True)
                    if rc != 0:
                        raise RuntimeError(f"ffmpeg encoder exited with {rc}")
                distributed.barrier()
            else:
                self._report_progress(0.9, "Encoding final video...")
                rc = enc.wait()
                if rc != 0:
                    raise RuntimeError(f"ffmpeg encoder exited with {rc}")
            self._report_progress(1.0, "Video processing complete!")
            return True
        except Exception as e:  # the reference's contract: report and return False
            self._report_progress(1.0, f"Error: {e}")
            print(f"Video processing error: {e}", file=sys.stderr)
            return False


# ---- raw-frame pipes (first slice) ----------------------------

def _decode_cmd(path: str, lo: int, hi: Optional[int]) -> List[str]:
    cmd = ["ffmpeg", "-v", "error", "-i", path]
    if lo or hi is not None:
        sel = f"trim=start_frame={lo}" + (f":end_frame={hi}" if hi is not None else "")
        cmd += ["-vf", sel + ",setpts=PTS-STARTPTS"]
    return cmd + ["-f", "rawvideo", "-pix_fmt", "rgb24", "-"]


def _iter_raw_frames(path: str, W: int, H: int, lo: int, hi: Optional[int], chunk: int):
    """Decoded frames [lo, hi) of the clip as u8 [n<=chunk, H, W, 3] arrays in pinned memory
    (two alternating buffers: the one yielded last stays valid until the next-but-one)."""
    frame_bytes = W * H * 3
    proc = subprocess.Popen(_decode_cmd(path, lo, hi), stdout=subprocess.PIPE,
                            stderr=subprocess.DEVNULL, bufsize=0)
    bufs = [pipeline.pinned_empty((chunk, H, W, 3)) for _ in range(2)]
    try:
        k = 0
        while True:
            buf = bufs[k & 1]
            mv = memoryview(buf).cast("B")
            got = 0
            want = chunk * frame_bytes
            while got < want:
                n = proc.stdout.readinto(mv[got:want])
                if not n:
                    break
                got += n
            nfr = got // frame_bytes
            if nfr:
                yield buf[:nfr]
            if got < want:
                break
            k += 1
        proc.stdout.close()
        rc = proc.wait()
        if rc != 0:
            raise RuntimeError(f"ffmpeg decoder exited with {rc}")
    finally:
        if proc.poll() is None:
            proc.kill()
        for b in bufs:
            pipeline.release_pinned(b)


def _read_raw_frames(path: str, W: int, H: int, lo: int, hi: int) -> np.ndarray:
    parts = [np.array(a) for a in _iter_raw_frames(path, W, H, lo, hi, max(1, hi - lo))]
    return np.concatenate(parts) if parts else np.empty((0, H, W, 3), np.uint8)


def _open_encoder(output_path: str, input_path: str, fps: float, ow: int, oh: int):
    """libx264 crf 18 yuv420p, audio and subtitles copied from the input (video_processor.py:
    361-382), fed with raw RGB frames on stdin."""
    cmd = ["ffmpeg", "-y", "-v", "error", "-f", "rawvideo", "-pix_fmt", "rgb24", "-s", f"{ow}x{oh}",
           "-framerate", f"{fps:.5f}", "-i", "-", "-i", input_path,
           "-map", "0:v:0", "-map", "1:a?", "-map", "1:s?",
           "-c:v", "libx264", "-preset", "medium", "-crf", "18", "-pix_fmt", "yuv420p",
           "-c:a", "copy", "-c:s", "copy", output_path]
    return subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.DEVNULL,
                            stderr=subprocess.DEVNULL)
