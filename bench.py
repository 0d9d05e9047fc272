#!/usr/bin/env python
"""bench.py -- headline benchmark of the dither_pie hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra]

Workload (BASELINE.json configs[1]): 3840x2160 RGB frames, error diffusion with the
Floyd-Steinberg + Atkinson + JJN kernels, 256-colour palette.  One "step" is one pass of the
three kernels over a batch of 128 synthetic 4K frames (3.2 GB, far larger than L2).
Metric: Mpixels/s (input pixels x dither passes per second), whole job over all ranks.
N > 1: one process per GPU (torchrun), every rank owns its own batch of frames (frames are
independent -> weak scaling, no data-path collective); time = max over ranks.

Keys besides the base contract:
  roofline      dominant kernel (k_diffuse_wave), algorithmic 6 B/pixel / measured launch time
                against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port (C restatement of the reference's numba loop) on the host cores
  e2e           the same metric through the C ABI with pinned HOST buffers (H2D + D2H timed)
  modes         the other BASELINE.json configs, one line each (device-resident, Mpx/s and
                fraction of the HBM roofline at 6 B/pixel)
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
# reference arm / cpu baseline: the oracle port on the host cores
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
        "config": workload_config(args.gpus, None),
        "cpu_baseline": {"value": val, "unit": "Mpx/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus, batch):
    return {"workload": "configs[1]: 3840x2160 RGB, error_diffusion floyd_steinberg+atkinson+jjn, "
                        "256-colour palette, serpentine=false",
            "frames_per_step_per_gpu": batch, "passes_per_frame": len(ED_VARIANTS),
            "palette": "first 256 unique rows of RandomState(2024).randint(0,256)",
            "frame": "synth.frame(2160,3840,seed) gradient + uniform noise [-16,16]",
            "l2_policy": "inputs larger than L2 (a batch of 128 4K frames is 3.2 GB in, 3x that out)",
            "parallelism": f"frame-sharded x{n_gpus}, no data-path collective"}


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

    from dither_pie_b200 import _capi, engine, synth
    from dither_pie_b200._capi import check, lib
    _capi.ensure_device(local)
    L = lib()
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    B = args.batch
    pal_rows = synth.random_palette(K_COLOURS)
    pal = engine.get_palette(pal_rows)
    # per-rank synthetic frames (seed depends on the rank so that ranks do different work)
    # The host side holds a block of HB frames in pinned memory; the device batch is that block
    # repeated (frames rolled per repeat would only change the bytes, not the work).  Every e2e
    # step still moves B frames in and 3 x B frames out over PCIe, block by block.
    HB = min(B, 32)
    assert B % HB == 0, "--batch must be a multiple of 32 (or smaller than 32)"
    base = np.stack([synth.frame(H4K, W4K, 1 + rank * 16 + t) for t in range(min(HB, 4))])
    host_in = _capi.PinnedArray((HB, H4K, W4K, 3), np.uint8)
    for t in range(HB):
        host_in.array[t] = base[t % base.shape[0]]
        if t >= base.shape[0]:
            host_in.array[t] = np.roll(host_in.array[t], 7 * t, axis=1)
    src = torch.empty((B, H4K, W4K, 3), dtype=torch.uint8, device=dev)
    dst = [torch.empty_like(src) for _ in ED_VARIANTS]
    blk = torch.from_numpy(host_in.array)
    for q in range(B // HB):
        src[q * HB:(q + 1) * HB].copy_(blk, non_blocking=False)
    plans = [engine.Plan("error_diffusion", {"variant": v}, H4K, W4K) for v in ED_VARIANTS]
    px_per_step = B * H4K * W4K * len(ED_VARIANTS)
    launches = 0

    def step():
        nonlocal launches
        for pl, d in zip(plans, dst):
            pl.run(pal, src.data_ptr(), B, d.data_ptr(), None, sp)
            launches += 2  # k_wave_init + k_diffuse_wave per call

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps * len(ED_VARIANTS))]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    k = 0
    for _ in range(args.steps):
        for pl, d in zip(plans, dst):
            ev[k][0].record(stream)
            pl.run(pal, src.data_ptr(), B, d.data_ptr(), None, sp)
            ev[k][1].record(stream)
            launches += 2
            k += 1
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * px_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- e2e: HOST buffers through the C ABI, copies inside the timed region ---------------
    # Three streams (copy-in, kernels, copy-out) and two device buffer sets: the H2D of step
    # n+1 and the D2H of kernel v overlap the kernels, which is how video_processor streams a
    # clip.  Every step still moves its whole input and all three results over PCIe.
    host_out = [_capi.PinnedArray((HB, H4K, W4K, 3), np.uint8) for _ in ED_VARIANTS]
    nbytes = B * H4K * W4K * 3
    blk_bytes = HB * H4K * W4K * 3
    s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    p_in, p_k, p_out = (C.c_void_p(x.cuda_stream) for x in (s_in, s_k, s_out))
    src2 = [src, torch.empty_like(src)]
    dst2 = [dst, [torch.empty_like(src) for _ in ED_VARIANTS]]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_k = [[torch.cuda.Event() for _ in ED_VARIANTS] for _ in range(2)]
    ev_out = [[torch.cuda.Event() for _ in ED_VARIANTS] for _ in range(2)]
    ev_src_free = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(nsteps):
        for n in range(nsteps):
            b = n & 1
            if n >= 2:
                s_in.wait_event(ev_src_free[b])          # kernels of step n-2 are done with src2[b]
            for q in range(B // HB):
                check(L.dp_memcpy_h2d(src2[b].data_ptr() + q * blk_bytes, host_in.ptr, blk_bytes, p_in),
                      "h2d")
            ev_in[b].record(s_in)
            s_k.wait_event(ev_in[b])
            for v, (pl, ho) in enumerate(zip(plans, host_out)):
                if n >= 2:
                    s_k.wait_event(ev_out[b][v])         # D2H of step n-2 has drained dst2[b][v]
                pl.run(pal, src2[b].data_ptr(), B, dst2[b][v].data_ptr(), None, p_k)
                ev_k[b][v].record(s_k)
                s_out.wait_event(ev_k[b][v])
                for q in range(B // HB):
                    check(L.dp_memcpy_d2h(ho.ptr, dst2[b][v].data_ptr() + q * blk_bytes, blk_bytes, p_out),
                          "d2h")
                ev_out[b][v].record(s_out)
            ev_src_free[b].record(s_k)
        for x in (s_in, s_k, s_out):
            x.synchronize()

    e2e_steps = max(3, min(args.steps, 6))
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * px_per_step * e2e_steps / float(te.item()) / 1e6
    # spot-check of the result that came back (first frame, first kernel) -- also keeps the
    # copies honest: the bytes must be palette colours
    chk = host_out[0].array[0, :4, :4].reshape(-1, 3)
    assert all(any((c == p).all() for p in pal_rows) for c in chk), "e2e output is not palette colours"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    avg_kernel_ms = sum(kernel_ms) / len(kernel_ms)
    alg_bytes = BYTES_PER_PX * B * H4K * W4K
    achieved = alg_bytes / (avg_kernel_ms * 1e-3) / 1e9
    # DRAM traffic per launch for 128 4K frames in the ncu --set full captures of this round:
    # 7.795 GB for k_diffuse_wave<floyd_steinberg>, 8.194 GB for <jjn> (the two ends of the step's
    # three launches; profiles/r1l_diffuse_wave_{fs,jjn}_K256_4k_x128.txt), i.e. 7.3-7.7 B/pixel
    # against 6 algorithmic -- the hand-off streams and the candidate table
    traffic = 0.5 * (7.794756e9 + 8.194386e9) / (128 * H4K * W4K) * (B * H4K * W4K)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                                  "mean of profiles/r1l_diffuse_wave_{fs,jjn}_K256_4k_x128.txt "
                                  "(scaled by frames)",
                "kernel": "k_diffuse_wave",
                "peak_source": peak_src, "avg_launch_ms": avg_kernel_ms,
                "note": "error diffusion is bounded by its per-pixel dependency chain (~1100 cycles "
                        "per wavefront step; 342 instructions per step for Floyd-Steinberg, 629 for "
                        "JJN, 58-63 % of the issue slots used) and by instruction issue, not by HBM; "
                        "see DESIGN.md"}

    line = {
        "metric": "Mpixels/s", "value": value, "unit": "Mpx/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world, B),
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_val, "unit": "Mpx/s", "h2d_bytes_per_step": nbytes,
                "d2h_bytes_per_step": nbytes * len(ED_VARIANTS), "steps": e2e_steps,
                "pipeline": "3 streams, double-buffered device batches; PCIe D2H-bound",
                "host_affinity": numa},
        "roofline": roofline,
    }
    if world == 1:
        # cpu baseline: bounded sample of the same workload on ALL host cores (undo the GPU-local
        # affinity first; worker threads inherit the mask when they are created)
        os.sched_setaffinity(0, orig_affinity)
        cores = len(orig_affinity) or 1
        crops = [np.ascontiguousarray(host_in.array[t % HB][:540, :960]) for t in range(cores)]
        t0 = time.perf_counter()
        px = cpu_reference_step(crops, pal_rows.astype(np.float32), cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": px / dt / 1e6, "unit": "Mpx/s", "cores": cores, "kind": "port",
            "sample": f"{cores} crops of 960x540 (one per core) x 3 kernels, {dt:.1f} s of wall time"}
        if not args.no_extra:
            try:
                line["modes"] = extra_modes(torch, engine, synth, sp, stream, peak)
            except Exception as e:  # never lose the headline line to an extra
                line["modes"] = {"error": str(e)[:200]}
    _emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()


def extra_modes(torch, engine, synth, sp, stream, peak):
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

    def entry(name, px, ms, alg_bytes=None, cpu=None, write_bytes=None):
        gbs = (alg_bytes if alg_bytes is not None else BYTES_PER_PX * px) / (ms * 1e-3) / 1e9
        out[name] = {"mpx_s": px / (ms * 1e-3) / 1e6, "ms": ms, "gb_s": gbs, "hbm_frac": gbs / peak,
                     "cpu_mpx_s": cpu}
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
        for (mode, params, K) in (("bayer", {"size": "8x8"}, 16), ("none", {}, 16),
                                  ("IGN", {}, 16), ("blue_noise", {}, 16),
                                  ("bayer", {"size": "8x8"}, 256), ("halftone", {}, 16)):
            if label == "4k" and mode in ("IGN", "blue_noise"):
                continue
            rows = pico if K == 16 else synth.random_palette(K)
            pal = engine.get_palette(rows)
            plan = engine.Plan(mode, params, h, w)
            ms = timed(lambda: plan.run(pal, src.data_ptr(), nf, dst.data_ptr(), None, sp))
            cpu = None
            if label == "1080p" and mode != "blue_noise":   # (its matrix takes seconds to generate)
                cpu = cpu_rate(frames[0], rows, mode, params)
            entry(f"{label}_{mode}_K{K}", nf * h * w, ms, cpu=cpu)
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
    # config 3: k-means Lloyd iteration over a full 4K frame, K=16 (3 B/pixel/iteration)
    from dither_pie_b200._capi import check, lib
    img = torch.from_numpy(synth.frame(2160, 3840, 2).reshape(-1, 3)).to(dev)
    n = img.shape[0]
    cent = torch.from_numpy(img[:: n // 16][:16].cpu().numpy().astype(np.float64)).to(dev)
    sums = torch.zeros(16 * 4, dtype=torch.int64, device=dev)
    ms = timed(lambda: check(lib().dp_kmeans_accumulate(img.data_ptr(), n, cent.data_ptr(), 16,
                                                        sums.data_ptr(), sp)))
    entry("4k_kmeans_lloyd_iter_K16", n, ms, 3.0 * n)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="4K frames per step per GPU")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
