#!/usr/bin/env python
"""bench.py -- benchmark of the dither_pie hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]

Headline workload (BASELINE.json configs[1]): 3840x2160 RGB frames, error diffusion with the
Floyd-Steinberg + Atkinson + JJN kernels, 256-colour palette.  One "step" is one pass of the three
kernels over a device-resident batch of 256 synthetic 4K frames (6.4 GB, far larger than L2); the
three kernels are independent and run on three streams.
Metric: Mpixels/s (input pixels x dither passes per second), whole job over all ranks.
N > 1: one process per GPU (torchrun), every rank owns its own batch of frames (frames are
independent -> weak scaling, no data-path collective); time = max over ranks.

Keys besides the base contract:
  roofline      dominant kernel (k_diffuse_wave), algorithmic 6 B/pixel / measured launch time
                against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port (C restatement of the reference's numba loop) on the host cores
  e2e           the same metric through the product's public API (pipeline.FramePipeline, the
                engine under VideoProcessor.process_frames) with pinned HOST arrays: every step
                copies its 128 input frames in and brings every result back; value = palette-index
                output (1 B/pixel, what north_star calls the output), `rgb_value` = colour bytes
  video         BASELINE configs[3] and [4] as STRONG-scaling jobs (600 x 1080p pixelize 270 ->
                blue noise / IGN -> x4; 300 x 4K Ostromoukhov / Sierra, 64 colours): the clip is
                split contiguously over the N ranks; device-resident time and end-to-end wall time
                through VideoProcessor.process_frames with host buffers, both max over ranks
  modes         the other BASELINE.json configs, one entry each (device-resident, Mpx/s and
                fraction of the HBM roofline), with the CPU baselines BASELINE.md section 3 asks for
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H4K, W4K = 2160, 3840
ED_VARIANTS = ("floyd_steinberg", "atkinson", "jjn")
K_COLOURS = 256
BYTES_PER_PX = 6.0  # 3 read + 3 written


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# reference arm / cpu baselines: the oracle on the host cores
# ------------------------------------------------------------------------------------------

def cpu_reference_step(frames_u8, palette, threads):
    """Each worker thread runs the C restatement of _error_diffusion_numba
    (dithering_lib.py:212-308; single-threaded per frame by construction) on one crop for the
    three kernels; ctypes releases the GIL.  Returns pixels x passes processed."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dither_oracle as O

    def work(img):
        flat = img.reshape(-1, 3).astype(np.float32)
        h, w, _ = img.shape
        for v in ED_VARIANTS:
            O.error_diffusion_indices(flat, palette, h, w, v, False)
        return h * w * len(ED_VARIANTS)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        return sum(ex.map(work, frames_u8))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is pure Python + numba, nothing to compile into oracle/_ref) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dither_pie_b200 import synth
    from oracle import dither_oracle as O
    O._lib()
    cores = os.cpu_count() or 1
    pal = synth.random_palette(K_COLOURS).astype(np.float32)
    ch, cw = 540, 960  # bounded sample: one 960x540 crop of a 4K frame per core and step
    crops = [np.ascontiguousarray(synth.frame(H4K, W4K, 1 + t)[:ch, :cw]) for t in range(min(cores, 4))]
    crops = [crops[i % len(crops)] for i in range(cores)]
    for _ in range(args.warmup):
        cpu_reference_step(crops[:cores], pal, cores)
    t0 = time.perf_counter()
    px = 0
    for _ in range(args.steps):
        px += cpu_reference_step(crops, pal, cores)
    dt = time.perf_counter() - t0
    val = px / dt / 1e6
    sample = (f"{cores} crops of {cw}x{ch} (one per core) of the 4K frame x 3 kernels per step; "
              "per-pixel cost is size-independent (O(K) palette scan)")
    line = {
        "impl": "reference", "metric": "Mpixels/s", "value": val, "unit": "Mpx/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the GPU arm's config: each step of this arm is a bounded SAMPLE of that workload
        # (cpu_baseline.sample says which)
        "config": workload_config(args.gpus, max(args.batch, args.dev_batch)),
        "cpu_baseline": {"value": val, "unit": "Mpx/s", "cores": cores, "kind": "port",
                         "sample": sample,
                         "note": "C port of the reference's numba loop on ALL cores; the reference's own "
                                 "loop is single-threaded (1.07 Mpx/s at 4K K=256, BASELINE.md)"},
        "e2e": {"value": val, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus, batch):
    return {"workload": "configs[1]: 3840x2160 RGB, error_diffusion floyd_steinberg+atkinson+jjn, "
                        "256-colour palette, serpentine=false",
            "frames_per_step_per_gpu": batch, "passes_per_frame": len(ED_VARIANTS),
            "palette": "first 256 unique rows of RandomState(2024).randint(0,256)",
            "frame": "synth.frame(2160,3840,seed) gradient + uniform noise [-16,16]",
            "l2_policy": f"inputs larger than L2 (a batch of {batch} 4K frames is "
                         f"{(batch or 0) * H4K * W4K * 3 / 1e9:.1f} GB in, 3x that out)",
            "streams": "the three variants of a step run on three streams (independent kernels over the "
                       "same batch; the next kernel's ramp-up fills the previous kernel's drain tail)",
            "parallelism": f"frame-sharded x{n_gpus}, no data-path collective"}


# ---- Pool.map video baseline: workers hold their frames, tasks are indices
_POOL = {}


def _pool_init(kind, pal_rows):
    from dither_pie_b200 import synth
    from oracle import dither_oracle as O
    _POOL["O"] = O
    _POOL["pal"] = np.asarray(pal_rows)
    if kind == "c4":
        _POOL["frames"] = [synth.frame(1080, 1920, 1000 + t) for t in range(4)]
        O.blue_noise_matrix(64, 42)                       # warm the cache (the reference's class cache)
    else:
        _POOL["frames"] = [np.ascontiguousarray(synth.frame(H4K, W4K, 2000)[:270, :480])]


def _pool_task(job):
    kind, mode, params, t = job
    O = _POOL["O"]
    f = _POOL["frames"][t % len(_POOL["frames"])]
    if kind == "c4":     # body of _process_single_frame minus PNG I/O (video_processor.py:443-462)
        out = O.final_resize(O.apply_dithering(O.pixelize_regular(f, 270), _POOL["pal"], mode, params), 4, True)
    else:
        out = O.apply_dithering(f, _POOL["pal"], mode, params)
    return int(out.shape[0])


def pool_video_baseline(kind, mode, params, pal_rows, procs, tasks):
    """Mpx/s of INPUT pixels for `tasks` frames mapped over a Pool of `procs` workers (spawned:
    the parent holds a CUDA context).  Pool start-up and frame synthesis are outside the timing;
    a first map warms the workers."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    px = (1080 * 1920) if kind == "c4" else (270 * 480)
    with ctx.Pool(procs, initializer=_pool_init, initargs=(kind, np.asarray(pal_rows))) as pool:
        pool.map(_pool_task, [(kind, mode, params, t) for t in range(procs)])
        t0 = time.perf_counter()
        pool.map(_pool_task, [(kind, mode, params, t) for t in range(tasks)], chunksize=1)
        dt = time.perf_counter() - t0
    return tasks * px / dt / 1e6, dt


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = statistics.median(sm)
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for j, nme in enumerate(names):
                    if any(r[5 + j].strip().lower().startswith("active") for r in rows if len(r) >= 9):
                        out["reasons"].append(nme)
                out["samples"] = len(sm)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

def _emit(line: dict, fd: int):
    """The one JSON line, written to the REAL stdout (fd saved before anything else could print)."""
    os.write(fd, (json.dumps(line) + "\n").encode())


class Arena:
    """One pinned host allocation carved into named arrays (pinning is slow: do it once)."""

    def __init__(self, nbytes):
        from dither_pie_b200 import _capi
        t0 = time.perf_counter()
        self.pa = _capi.PinnedArray((nbytes,), np.uint8)
        self.alloc_s = time.perf_counter() - t0
        self.nbytes = nbytes

    def view(self, offset, shape):
        n = int(np.prod(shape))
        assert offset + n <= self.nbytes, (offset, n, self.nbytes)
        return self.pa.array[offset:offset + n].reshape(shape)


def run_gpu(args):
    # Libraries print to stdout (NCCL's version banner at the first collective): keep the process's
    # stdout for the one JSON line and send everything else to stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # Keep this rank (and the pinned host buffers it first-touches) on the NUMA node of its GPU:
    # with 8 ranks the e2e leg is bound by host memory and PCIe, not by the kernels.
    numa = "unpinned"
    orig_affinity = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local)
            hnd = pynvml.nvmlDeviceGetHandleByPciBusId(
                f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(hnd)
        numa = f"cpu affinity of GPU {local}: {len(os.sched_getaffinity(0))} cores"
    except Exception as e:   # best effort
        numa = f"unpinned ({type(e).__name__})"
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from dither_pie_b200 import _capi, engine, pipeline, synth
    _capi.ensure_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    pal_rows = synth.random_palette(K_COLOURS)
    pal = engine.get_palette(pal_rows)
    # ---- pinned host arena: [inputs | results]; every e2e / video leg carves its arrays from it
    frame4k = H4K * W4K * 3
    in_bytes = max(B * frame4k, 300 // world * frame4k + frame4k, 600 // world * 1080 * 1920 * 3 + 1080 * 1920 * 3)
    out_bytes = max(B * frame4k * (len(ED_VARIANTS) if not args.no_rgb_e2e else 1), in_bytes)
    arena = Arena(in_bytes + out_bytes)
    host_in = arena.view(0, (B, H4K, W4K, 3))
    base = np.stack([synth.frame(H4K, W4K, 1 + rank * 16 + t) for t in range(4)])
    for t in range(B):       # per-rank synthetic frames: 4 seeds, the rest rolled copies (new bytes)
        host_in[t] = base[t % 4] if t < 4 else np.roll(base[t % 4], 7 * t, axis=1)
    # device-resident batch: DB frames (the wavefront kernel loses its ramp-up and drain tail once
    # per launch, so larger batches amortise them: 128 -> 256 frames is worth ~10 %); the first B
    # are the host frames of the e2e legs, the rest rolled copies (new bytes)
    DB = max(B, args.dev_batch)
    src = torch.empty((DB, H4K, W4K, 3), dtype=torch.uint8, device=dev)
    src[:B].copy_(torch.from_numpy(host_in))
    for t0 in range(B, DB, B):
        n = min(B, DB - t0)
        src[t0:t0 + n].copy_(torch.roll(src[:n], shifts=11 * (t0 // B), dims=2))
    dst = [torch.empty_like(src) for _ in ED_VARIANTS]
    plans = [engine.Plan("error_diffusion", {"variant": v}, H4K, W4K) for v in ED_VARIANTS]
    px_per_step = B * H4K * W4K * len(ED_VARIANTS)
    dev_px_per_step = DB * H4K * W4K * len(ED_VARIANTS)
    launches = 0
    # one stream per variant: the three kernels of a step are independent (same input, own output);
    # the longest (JJN) is launched first
    vstreams = [torch.cuda.Stream(device=dev) for _ in ED_VARIANTS]
    vsp = [C.c_void_p(v.cuda_stream) for v in vstreams]
    order = sorted(range(len(ED_VARIANTS)), key=lambda i: ED_VARIANTS[i] != "jjn")
    fork = torch.cuda.Event()
    joins = [torch.cuda.Event() for _ in ED_VARIANTS]

    def step():
        nonlocal launches
        fork.record(stream)
        for i in order:
            vstreams[i].wait_event(fork)
            plans[i].run(pal, src.data_ptr(), DB, dst[i].data_ptr(), None, vsp[i])
            joins[i].record(vstreams[i])
            launches += 2  # k_wave_init + k_diffuse_wave per call
        for i in order:
            stream.wait_event(joins[i])

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world * dev_px_per_step * args.steps / (ms_total * 1e-3) / 1e6
    # the same launches one after the other on ONE stream (not part of `value`): the duration of
    # each kernel by itself, for the roofline line and for comparison with the ncu launch list
    iso_steps = max(2, min(args.steps, 4))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(iso_steps * len(ED_VARIANTS))]
    k = 0
    for _ in range(iso_steps):
        for pl, d in zip(plans, dst):
            ev[k][0].record(stream)
            pl.run(pal, src.data_ptr(), DB, d.data_ptr(), None, sp)
            ev[k][1].record(stream)
            k += 1
    barrier()
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    per_variant_ms = {v: statistics.mean(kernel_ms[i::len(ED_VARIANTS)]) for i, v in enumerate(ED_VARIANTS)}
    del dst

    # ---- e2e: pinned HOST arrays through pipeline.FramePipeline (the engine under
    # VideoProcessor.process_frames): every step submits its 128 input frames and gets all three
    # results back; copy-in / kernels / copy-out overlap across batches and steps.
    nbytes = B * frame4k
    e2e_steps = max(3, min(args.steps, 6))
    PB = min(B, args.pipe_batch)       # frames per device batch inside the pipeline

    def e2e_leg(output, steps):
        if output == "index":
            outs = [arena.view(in_bytes + v * B * H4K * W4K, (B, H4K, W4K)) for v in range(len(ED_VARIANTS))]
            kw = {"out_idx": outs}
        else:
            outs = [arena.view(in_bytes + v * nbytes, (B, H4K, W4K, 3)) for v in range(len(ED_VARIANTS))]
            kw = {"out_rgb": outs}
        with pipeline.FramePipeline(plans, pal, PB, output) as pipe:
            pipe.run(host_in, **kw)                       # warm-up (buffers, workspaces)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                pipe.submit(host_in, **kw)
            pipe.flush()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            stats = dict(pipe.stats)
        return outs, world * px_per_step * steps / dt / 1e6, stats

    outs, e2e_val, st_idx = e2e_leg("index", e2e_steps)
    # the result that came back must be the kernels' answer: compare one frame per variant with a
    # direct device-resident call
    chk = torch.empty((1, H4K, W4K), dtype=torch.uint8, device=dev)
    for v, pl in enumerate(plans):
        pl.run(pal, src.data_ptr() + 5 * frame4k, 1, None, chk.data_ptr(), sp)
        torch.cuda.synchronize()
        assert np.array_equal(chk[0].cpu().numpy(), outs[v][5]), "e2e result differs from the direct call"
    e2e = {"value": e2e_val, "unit": "Mpx/s", "h2d_bytes_per_step": nbytes,
           "d2h_bytes_per_step": B * H4K * W4K * len(ED_VARIANTS), "steps": e2e_steps,
           "api": "dither_pie_b200.pipeline.FramePipeline.submit/flush (pinned host arrays, index-plane output)",
           "output": "palette-index plane, 1 B/pixel per variant",
           "pipeline": f"copy-in + one kernel stream per variant + copy-out, {PB}-frame device batches, "
                       f"{pipeline.FramePipeline.SLOTS} buffer sets",
           "h2d_gbs_per_rank": st_idx["h2d_gbs"], "d2h_gbs_per_rank": st_idx["d2h_gbs"],
           "host_affinity": numa, "pinned_arena_gb": arena.nbytes / 1e9, "pinned_alloc_s": arena.alloc_s}
    if not args.no_rgb_e2e:
        _, rgb_val, st_rgb = e2e_leg("rgb", 2)
        e2e.update({"rgb_value": rgb_val, "rgb_d2h_bytes_per_step": nbytes * len(ED_VARIANTS),
                    "rgb_d2h_gbs_per_rank": st_rgb["d2h_gbs"],
                    "rgb_note": "same call with colour-byte output (3 B/pixel per variant): PCIe D2H-bound"})
    del src
    torch.cuda.empty_cache()

    # ---- video configs as strong-scaling jobs (all ranks)
    video = None
    if not args.no_video:
        try:
            video = video_lines(torch, dist, engine, synth, arena, in_bytes, rank, world, dev, sp, stream,
                                barrier, max_over_ranks)
        except Exception as e:
            video = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- the one collective of the design: pixel-sharded k-means, NCCL all-reduce of the K x 4
    # integer sums on the kernel stream (all ranks)
    kmeans_line = None
    if not args.no_video:
        try:
            kmeans_line = kmeans_sharded_line(torch, dist, synth, rank, world, dev, max_over_ranks, barrier)
        except Exception as e:
            kmeans_line = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # the three launches of a step overlap on three streams: the average launch duration inside
    # the timed region is the step time / 3 (each kernel by itself: launch_ms_isolated)
    avg_kernel_ms = ms_total / args.steps / len(ED_VARIANTS)
    alg_bytes = BYTES_PER_PX * DB * H4K * W4K
    achieved = alg_bytes / (avg_kernel_ms * 1e-3) / 1e9
    # DRAM traffic per launch for 128 4K frames in the ncu --set full captures (profiles/):
    # 7.617 GB for k_diffuse_wave<floyd_steinberg>, 7.761 GB for <jjn> (the two ends of the step's
    # three launches), i.e. 7.2-7.3 B/pixel against 6 algorithmic -- hand-off streams + table
    traffic = 0.5 * (7.616983e9 + 7.761e9) / (128 * H4K * W4K) * (DB * H4K * W4K)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                                  "mean of profiles/r2k_diffuse_wave_{fs,jjn}_K256_4k_x128.txt "
                                  "(scaled by frames)",
                "kernel": "k_diffuse_wave",
                "peak_source": peak_src, "avg_launch_ms": avg_kernel_ms,
                "avg_launch_note": "the step's three launches overlap on three streams: step time / 3, "
                                   "measured with CUDA events on the stream that forks and joins them",
                "launch_ms_isolated": per_variant_ms,
                "launch_isolated_note": "each kernel by itself, one stream, CUDA events around the launch "
                                        "(a separate pass after the timed region)",
                "note": "error diffusion is bounded by its per-pixel dependency chain (~1100 cycles "
                        "per wavefront step) and by instruction issue, not by HBM; see DESIGN.md"}

    line = {
        "metric": "Mpixels/s", "value": value, "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world, DB),
        "clocks": clocks, "gpu_launches": launches,
        "e2e": e2e, "roofline": roofline,
    }
    if video is not None:
        line["video"] = video
    if kmeans_line is not None:
        line["kmeans_sharded"] = kmeans_line
    if world == 1:
        # cpu baseline: bounded sample of the same workload on ALL host cores (undo the GPU-local
        # affinity first; worker threads inherit the mask when they are created)
        os.sched_setaffinity(0, orig_affinity)
        cores = len(orig_affinity) or 1
        crops = [np.ascontiguousarray(host_in[t % B][:540, :960]) for t in range(cores)]
        t0 = time.perf_counter()
        px = cpu_reference_step(crops, pal_rows.astype(np.float32), cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": px / dt / 1e6, "unit": "Mpx/s", "cores": cores, "kind": "port",
            "sample": f"{cores} crops of 960x540 (one per core) x 3 kernels, {dt:.1f} s of wall time",
            "note": "C port of the reference's numba loop on ALL cores (the reference's own loop is "
                    "single-threaded: 1.07 Mpx/s at 4K K=256, BASELINE.md)"}
        if not args.no_extra:
            try:
                line["modes"] = extra_modes(torch, engine, synth, sp, stream, peak, cores)
            except Exception as e:  # never lose the headline line to an extra
                line["modes"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    _emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()


def kmeans_sharded_line(torch, dist, synth, rank, world, dev, max_over_ranks, barrier):
    """BASELINE configs[2], throughput mode, as a strong-scaling job: Lloyd iterations (K=16) over
    the 8.29 Mpx 4K frame, pixels sharded contiguously over the ranks, stop test on the device
    (one persistent launch runs the whole loop).  Exchange of the integer sums between the ranks:
    "p2p" -- block 0 pushes them into every rank's inbox over NVLink after the grid barrier that
    ends the assignment pass, every block waits for the flags (dp_kmeans_lloyd_p2p) -- and, for
    comparison, a launch pair per iteration with ncclAllReduce on the kernel stream.  Wall clock of the synchronous call (workspace allocation, centre upload and
    the final read-back included), max over ranks, best of 3, at 20 and at 100 iterations; the
    marginal cost per iteration is the difference.  The SHA-256 of the final centres must be the
    same for every N and both exchanges (integer sums -> shard-count invariant)."""
    import hashlib
    from dither_pie_b200 import _capi, distributed as D, kmeans as KM
    img = synth.frame(H4K, W4K, 2).reshape(-1, 3)
    init = img[np.random.RandomState(0).choice(len(img), 16, replace=False)].astype(np.float64)
    lo, hi = D.shard_pixels(len(img), rank, world)
    shard = _capi.DeviceBuffer((hi - lo) * 3).upload(np.ascontiguousarray(img[lo:hi]))

    def timed(iters, kw):
        KM.lloyd_device(shard.ptr, hi - lo, init, -1.0, 2, **kw())          # warm-up
        best, res = None, None
        for _ in range(3):
            args = kw()
            barrier()
            t0 = time.perf_counter()
            res = KM.lloyd_device(shard.ptr, hi - lo, init, -1.0, iters, check_every=iters, **args)
            dt = max_over_ranks(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        return best, res

    def leg(kw):
        t20, res = timed(20, kw)
        t100, _ = timed(100, kw)
        return {"ms_20_iterations": t20 * 1e3, "ms_100_iterations": t100 * 1e3,
                "us_per_iteration": (t100 - t20) / 80 * 1e6, "fixed_ms": (t20 - (t100 - t20) / 4) * 1e3,
                "mpx_s": len(img) * 80 / (t100 - t20) / 1e6, "tied_samples": int(res.ties),
                "centres_sha256_16": hashlib.sha256(np.ascontiguousarray(res[0]).tobytes()).hexdigest()[:16]}

    out = {"n_gpus": world, "scaling": "strong", "pixels": int(len(img)), "K": 16}
    if world == 1:
        out.update(leg(lambda: {}))
        out["exchange"] = "none"
    else:
        out.update(leg(lambda: {"p2p": D.p2p_exchange()}))
        out["exchange"] = ("peer memory inside the persistent loop kernel: after the grid barrier that ends the "
                           "assignment pass block 0 stores the 65 u64 sums into every rank's inbox over NVLink "
                           "(cudaIpc mappings) and releases the flags; every block waits for the flags and adds "
                           "the slots up")
        comm = D.nccl_comm()
        out["nccl"] = leg(lambda: {"comm": comm})
        out["nccl"]["exchange"] = "ncclAllReduce(u64 x 65) per iteration on the kernel stream"
    shard.free()
    return out


def video_lines(torch, dist, engine, synth, arena, in_bytes, rank, world, dev, sp, stream, barrier, max_over_ranks):
    """BASELINE configs[3] / [4] as strong-scaling jobs: a fixed clip split contiguously over the
    ranks.  `device_ms`: frames resident in HBM, one launch per rank over its shard, CUDA events,
    max over ranks.  `e2e_ms`: VideoProcessor.process_frames on pinned host arrays (copies inside
    the timed region), wall clock between barriers, max over ranks."""
    import dither_pie_b200 as dp
    from dither_pie_b200.dithering_lib import DitherMode
    from dither_pie_b200.video_processor import VideoProcessor, shard_frames
    out = {}

    def timed_dev(fn, reps):
        fn()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ms = max_over_ranks(a.elapsed_time(b))
            best = ms if best is None else min(best, ms)
        return best

    def timed_wall(fn, reps):
        fn()
        best = None
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fn()
            barrier()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            best = ms if best is None else min(best, ms)
        return best

    # ---- config 4: 600 x 1080p, pixelize 270 -> blue noise / IGN -> x4, K=16 from frame 0
    NF, H, W = 600, 1080, 1920
    lo, hi = shard_frames(NF, rank, world)
    n = hi - lo
    base = [synth.frame(H, W, 1000 + t) for t in range(8)]     # 8 seeds, repeated cyclically
    h_in = arena.view(0, (n, H, W, 3))
    for t in range(n):
        h_in[t] = base[(lo + t) % 8]
    h_out = arena.view(in_bytes, (n, H, W, 3))
    h_idx = arena.view(in_bytes, (n, 270, 480))
    d_in = torch.from_numpy(h_in).to(dev)
    d_out = torch.empty_like(d_in)
    pal16 = None
    for mode, params, tag in ((DitherMode.BLUE_NOISE, {"size": 64, "seed": 42}, "blue_noise"),
                              (DitherMode.INTERLEAVED_GRADIENT_NOISE, {"scale": 1.0, "seed": 0}, "IGN")):
        d = dp.ImageDitherer(num_colors=16, dither_mode=mode, palette=pal16, dither_params=params)
        vp = VideoProcessor()
        vp._setup(base[0][None], d, ("regular", 270))     # palette from frame 0 (median cut), every rank
        pal16 = d.palette
        plan = engine.Plan(mode.value, params, 270, 480, (H, W), 4)
        palh = engine.get_palette(pal16)
        dev_ms = timed_dev(lambda: plan.run(palh, d_in.data_ptr(), n, d_out.data_ptr(), None, sp), 5)
        e2e_ms = timed_wall(lambda: vp.process_frames(h_in, d, ("regular", 270), final_resize_multiplier=4,
                                                      out=h_out, shard=False), 3)
        st = dict(vp.last_stats)
        idx_ms = timed_wall(lambda: vp.process_frames(h_in, d, ("regular", 270), final_resize_multiplier=4,
                                                      out=h_idx, shard=False, output="index"), 3)
        out[f"config4_{tag}"] = {
            "frames": NF, "frames_per_rank": n, "n_gpus": world, "scaling": "strong",
            "device_ms": dev_ms, "device_mpx_s": NF * H * W / dev_ms / 1e3,
            "e2e_ms": e2e_ms, "e2e_mpx_s": NF * H * W / e2e_ms / 1e3,
            "e2e_index_ms": idx_ms, "e2e_index_mpx_s": NF * H * W / idx_ms / 1e3,
            "h2d_gbs_rank0": st.get("h2d_gbs"), "d2h_gbs_rank0": st.get("d2h_gbs"),
            "api": "VideoProcessor.process_frames(pinned frames, ('regular', 270), final_resize_multiplier=4)"}
        vp.close()
    del d_in, d_out
    torch.cuda.empty_cache()

    # ---- config 5: 300 x 4K, Ostromoukhov and Sierra, K=64
    NF, H, W = 300, H4K, W4K
    lo, hi = shard_frames(NF, rank, world)
    n = hi - lo
    base = [synth.frame(H, W, 2000 + t) for t in range(4)]
    h_in = arena.view(0, (n, H, W, 3))
    for t in range(n):
        h_in[t] = base[(lo + t) % 4] if (lo + t) < 4 else np.roll(base[(lo + t) % 4], 5 * (lo + t), axis=1)
    h_out = arena.view(in_bytes, (n, H, W, 3))
    h_idx = arena.view(in_bytes, (n, H, W))
    d_in = torch.from_numpy(h_in).to(dev)
    d_out = torch.empty_like(d_in)
    pal64 = [tuple(int(v) for v in r) for r in synth.random_palette(64)]
    palh = engine.get_palette(pal64)
    for mode, params, tag in ((DitherMode.OSTROMOUKHOV, {}, "ostromoukhov"),
                              (DitherMode.ERROR_DIFFUSION, {"variant": "sierra"}, "sierra")):
        d = dp.ImageDitherer(num_colors=64, dither_mode=mode, palette=pal64, dither_params=params)
        vp = VideoProcessor()
        plan = engine.Plan(mode.value, params, H, W)
        dev_ms = timed_dev(lambda: plan.run(palh, d_in.data_ptr(), n, d_out.data_ptr(), None, sp), 2)
        e2e_ms = timed_wall(lambda: vp.process_frames(h_in, d, out=h_out, shard=False), 2)
        st = dict(vp.last_stats)
        idx_ms = timed_wall(lambda: vp.process_frames(h_in, d, out=h_idx, shard=False, output="index"), 2)
        out[f"config5_{tag}"] = {
            "frames": NF, "frames_per_rank": n, "n_gpus": world, "scaling": "strong",
            "device_ms": dev_ms, "device_mpx_s": NF * H * W / dev_ms / 1e3,
            "e2e_ms": e2e_ms, "e2e_mpx_s": NF * H * W / e2e_ms / 1e3,
            "e2e_index_ms": idx_ms, "e2e_index_mpx_s": NF * H * W / idx_ms / 1e3,
            "h2d_gbs_rank0": st.get("h2d_gbs"), "d2h_gbs_rank0": st.get("d2h_gbs"),
            "api": "VideoProcessor.process_frames(pinned frames)"}
        vp.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    out["note"] = ("strong scaling: fixed clip, contiguous frame shards, no data-path collective; times are "
                   "best of 2-5 repetitions, max over ranks; synthetic frames repeat 8 (1080p) / 4 (4K, rolled) "
                   "seeds cyclically")
    return out


def extra_modes(torch, engine, synth, sp, stream, peak, cores):
    """The other BASELINE.json configs, device-resident, batched (>= L2), 5 timed repetitions.
    Each entry also carries ``cpu_mpx_s``: the oracle (the reference's algorithm: numpy + scipy
    KD-tree with all host threads, C port of the numba loop) on a bounded sample of the same
    mode -- a baseline, not a target."""
    from oracle import dither_oracle as O
    dev = torch.device("cuda", torch.cuda.current_device())
    out = {}

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def cpu_rate(img, pal_rows, mode, params):
        """Mpx/s of the oracle on one (cropped) frame; the first call warms caches."""
        try:
            O.apply_dithering(img[:64, :64], pal_rows, mode, params)
            t0 = time.perf_counter()
            O.apply_dithering(img, pal_rows, mode, params)
            return img.shape[0] * img.shape[1] / (time.perf_counter() - t0) / 1e6
        except Exception:
            return None

    def entry(name, px, ms, alg_bytes=None, cpu=None, write_bytes=None, **more):
        gbs = (alg_bytes if alg_bytes is not None else BYTES_PER_PX * px) / (ms * 1e-3) / 1e9
        out[name] = {"mpx_s": px / (ms * 1e-3) / 1e6, "ms": ms, "gb_s": gbs, "hbm_frac": gbs / peak,
                     "cpu_mpx_s": cpu}
        out[name].update(more)
        if write_bytes is not None:
            # write-dominated path: also against the WRITE-ONLY bandwidth measured in this run
            out[name]["write_gb_s"] = write_bytes / (ms * 1e-3) / 1e9
            out[name]["write_only_peak_gb_s"] = write_peak
            out[name]["write_frac"] = out[name]["write_gb_s"] / write_peak

    # write-only HBM bandwidth (fill of 1.6 GB): the bound of the x4 up-scaling video path, whose
    # output is 16 times its input -- a fill reaches ~3.9 TB/s on B200, a copy 6.5 TB/s (r+w)
    fill = torch.empty(1600 * 1000 * 1000, dtype=torch.uint8, device=dev)
    write_peak = fill.numel() / (timed(lambda: fill.fill_(1), 10) * 1e-3) / 1e9
    del fill

    pico = synth.hex_palette(synth.PICO8)
    for (label, h, w, nf) in (("1080p", 1080, 1920, 64), ("4k", 2160, 3840, 16)):
        frames = np.stack([synth.frame(h, w, t) for t in range(2)])
        src = torch.from_numpy(np.concatenate([frames] * (nf // 2))).to(dev)
        dst = torch.empty_like(src)
        idx = torch.empty((nf, h, w), dtype=torch.uint8, device=dev)
        for (mode, params, K) in (("bayer", {"size": "8x8"}, 16), ("none", {}, 16),
                                  ("IGN", {}, 16), ("blue_noise", {}, 16),
                                  ("bayer", {"size": "8x8"}, 256), ("none", {}, 256), ("halftone", {}, 16)):
            if label == "4k" and mode in ("IGN", "blue_noise"):
                continue
            rows = pico if K == 16 else synth.random_palette(K)
            pal = engine.get_palette(rows)
            plan = engine.Plan(mode, params, h, w)
            ms = timed(lambda: plan.run(pal, src.data_ptr(), nf, dst.data_ptr(), None, sp))
            ms_idx = timed(lambda: plan.run(pal, src.data_ptr(), nf, None, idx.data_ptr(), sp))
            cpu = None
            if label == "1080p" and mode != "blue_noise":   # (its matrix takes seconds to generate)
                cpu = cpu_rate(frames[0], rows, mode, params)
            entry(f"{label}_{mode}_K{K}", nf * h * w, ms, cpu=cpu, index_only_ms=ms_idx,
                  index_only_hbm_frac=4.0 * nf * h * w / (ms_idx * 1e-3) / 1e9 / peak)
        del idx
        pal64_rows = synth.random_palette(64)
        pal64 = engine.get_palette(pal64_rows)
        if label == "1080p":
            plan = engine.Plan("error_diffusion", {"variant": "sierra"}, h, w)
            ms = timed(lambda: plan.run(pal64, src.data_ptr(), nf, dst.data_ptr(), None, sp), 3)
            entry(f"{label}_ed_sierra_K64", nf * h * w, ms,
                  cpu=cpu_rate(frames[0][:540, :960], pal64_rows, "error_diffusion", {"variant": "sierra"}))
            # config 4: pixelize 1080p -> 480x270, dither, x4 up-scale, fused
            pal16 = engine.get_palette(pico)
            for mode in ("blue_noise", "IGN"):
                plan = engine.Plan(mode, {}, 270, 480, (h, w), 4)
                ms = timed(lambda: plan.run(pal16, src.data_ptr(), nf, dst.data_ptr(), None, sp))
                entry(f"video1080p_pixelize270_{mode}_x4", nf * h * w, ms,
                      nf * (3 * 480 * 270 + 3 * 1920 * 1080), write_bytes=nf * 3 * 1920 * 1080)
        else:
            # config 2, single image: latency of ONE 4K frame per kernel (the reference: 7.7 s)
            pal256_rows = synth.random_palette(256)
            pal256 = engine.get_palette(pal256_rows)
            for v in ED_VARIANTS:
                plan = engine.Plan("error_diffusion", {"variant": v}, h, w)
                ms = timed(lambda: plan.run(pal256, src.data_ptr(), 1, dst.data_ptr(), None, sp), 5)
                entry(f"4k_single_image_ed_{v}_K256", h * w, ms, reference_numba_s=7.7)
            # serpentine is serial per frame (one warp walks the frame): parallel over frames only
            plan = engine.Plan("error_diffusion", {"variant": "floyd_steinberg", "serpentine": "true"}, 1080, 1920)
            ms = timed(lambda: plan.run(pal256, src.data_ptr(), 1, dst.data_ptr(), None, sp), 1)
            entry("1080p_single_image_ed_fs_serpentine_K256", 1080 * 1920, ms)
            # config 5: 4K, 64 colours, Ostromoukhov and Sierra; 300 frames over 8 GPUs = 38 frames
            # per GPU (frames shard over GPUs), which is also what saturates the wavefront kernel
            nf5 = 38
            src5 = torch.from_numpy(np.concatenate([frames] * (nf5 // 2))).to(dev)
            dst5 = torch.empty_like(src5)
            for (mode, params, tag) in (("ostromoukhov", {}, "ostromoukhov"),
                                        ("error_diffusion", {"variant": "sierra"}, "ed_sierra"),
                                        ("hybrid", {}, "hybrid"), ("perceptual", {}, "perceptual"),
                                        ("adaptive_variance", {"var_threshold": 60.0}, "adaptive_variance")):
                plan = engine.Plan(mode, params, h, w)
                ms = timed(lambda: plan.run(pal64, src5.data_ptr(), nf5, dst5.data_ptr(), None, sp), 3)
                crop = frames[0][:540, :960]
                entry(f"{label}_{tag}_K64_x{nf5}", nf5 * h * w, ms, cpu=cpu_rate(crop, pal64_rows, mode, params))
            del src5, dst5
        del src, dst
    torch.cuda.empty_cache()

    # ---- config 3: k-means Lloyd iterations, K=16, 3 B/pixel/iteration.  16 DISTINCT 4K frames
    # (398 MB, larger than L2) as one pixel array; then the full-image loop on the seed-2 frame.
    from dither_pie_b200 import kmeans as KM
    from dither_pie_b200._capi import check, lib
    frames = np.concatenate([synth.frame(2160, 3840, 2 + t).reshape(-1, 3) for t in range(16)])
    img = torch.from_numpy(frames).to(dev)
    n = img.shape[0]
    rs = np.random.RandomState(0)
    init = frames[rs.choice(n, 16, replace=False)].astype(np.float64)
    cent = torch.from_numpy(init).to(dev)
    sums = torch.zeros(16 * 4 + 1, dtype=torch.int64, device=dev)
    ms = timed(lambda: check(lib().dp_kmeans_accumulate(img.data_ptr(), n, cent.data_ptr(), 16,
                                                        sums.data_ptr(), sp)))
    entry("kmeans_assign_pass_first_iteration_K16_16x4k", n, ms, 3.0 * n,
          note="one assignment pass (grid build + k_kmeans_accum16) over 133 Mpx = 398 MB (> L2) with the INITIAL "
               "centres (16 random pixels): the worst case, whole image regions lie in boxes that keep more than "
               "four candidate centres")
    cent5 = torch.from_numpy(np.ascontiguousarray(KM.lloyd_device(img.data_ptr(), n, init, -1.0, 5)[0])).to(dev)
    ms = timed(lambda: check(lib().dp_kmeans_accumulate(img.data_ptr(), n, cent5.data_ptr(), 16,
                                                        sums.data_ptr(), sp)))
    entry("kmeans_assign_pass_K16_16x4k", n, ms, 3.0 * n,
          note="the same pass with the centres after 5 Lloyd iterations (what every later iteration looks like)")
    n1 = 2160 * 3840
    KM.lloyd_device(img.data_ptr(), n1, init, -1.0, 2)          # first-call set-up (function attributes)
    loop_ms = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = KM.lloyd_device(img.data_ptr(), n1, init, -1.0, 100)
        dt = (time.perf_counter() - t0) * 1e3
        loop_ms = dt if loop_ms is None else min(loop_ms, dt)
    # sklearn on the same full array: seconds per Lloyd iteration
    sk_iter_s = None
    try:
        from sklearn.cluster import KMeans
        X = frames[:n1].astype(np.float64)
        t0 = time.perf_counter()
        km = KMeans(n_clusters=16, init=init, n_init=1, max_iter=3, tol=0.0, algorithm="lloyd").fit(X)
        sk_iter_s = (time.perf_counter() - t0) / max(int(km.n_iter_), 1)
    except Exception:
        pass
    entry("kmeans_lloyd_loop_K16_4k", n1 * res[1], loop_ms, 3.0 * n1 * res[1],
          cpu=(n1 / sk_iter_s / 1e6) if sk_iter_s else None,
          iterations=res[1], ms_per_iteration=loop_ms / max(res[1], 1), tied_samples=res.ties,
          sklearn_s_per_iteration=sk_iter_s, cpu_cores=cores,
          note="dp_kmeans_lloyd: 100 iterations over the 8.29 Mpx frame in ONE persistent launch (grid barriers "
               "between the centre update and the assignment pass, stop test on the device), wall clock of the "
               "synchronous call incl. workspace allocation and read-back, best of 3; cpu = "
               "sklearn.cluster.KMeans (the call the reference makes) on the full array")
    del img
    torch.cuda.empty_cache()

    # ---- CPU baselines of the video configs: Pool(P).map over in-memory
    # frames, P = all cores and P = the reference default min(4, cores - 1)
    try:
        pal16_rows = [tuple(int(v) for v in r) for r in pico]
        for P in sorted({cores, min(4, max(1, cores - 1))}):
            for mode, params in (("blue_noise", {"size": 64, "seed": 42}), ("IGN", {"scale": 1.0, "seed": 0})):
                mpx, dt = pool_video_baseline("c4", mode, params, pal16_rows, P, 4 * P)
                out[f"cpu_pool{P}_config4_{mode}"] = {
                    "cpu_mpx_s": mpx, "workers": P, "frames": 4 * P, "seconds": dt,
                    "note": "oracle pixelize->dither->x4 per 1080p frame, multiprocessing.Pool.map, input Mpx/s"}
        P = cores
        mpx, dt = pool_video_baseline("c5", "ostromoukhov", {}, synth.random_palette(64), P, P)
        out[f"cpu_pool{P}_config5_ostromoukhov"] = {
            "cpu_mpx_s": mpx, "workers": P, "frames": P, "seconds": dt, "kind": "port",
            "note": "C port of the live Ostromoukhov path on one 480x270 crop per core, extrapolated linearly "
                    "in pixels; the reference's own path is pure Python at ~0.014 Mpx/s per core (BASELINE.md)"}
    except Exception as e:
        out["cpu_pool_error"] = f"{type(e).__name__}: {e}"[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="4K host frames per e2e step per GPU")
    ap.add_argument("--dev-batch", type=int, default=256,
                    help="4K frames per device-resident step per GPU (>= --batch)")
    ap.add_argument("--pipe-batch", type=int, default=64, help="frames per device batch of the e2e pipeline")
    ap.add_argument("--no-extra", action="store_true", help="skip the per-mode entries")
    ap.add_argument("--no-video", action="store_true", help="skip the strong-scaling video configs")
    ap.add_argument("--no-rgb-e2e", action="store_true", help="skip the colour-byte variant of the e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
