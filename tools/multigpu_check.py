"""torchrun --nproc-per-node N tools/multigpu_check.py

Checks the N > 1 paths on real GPUs: (1) pixel-sharded k-means (NCCL all-reduce of the integer
sums) gives the centres of the single-GPU run; (2) frame-sharded dithering gives the frames of
the single-GPU run; prints per-rank device timings (max over ranks is the job time)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dither_pie_b200 import _capi, distributed as D, engine, kmeans, synth  # noqa: E402


def main():
    rank, world = D.init_process_group()
    _capi.ensure_device(int(os.environ.get("LOCAL_RANK", "0")))
    # ---- k-means, 4K frame, K=16: pixel-sharded Lloyd loop, NCCL all-reduce on the kernel stream
    img = synth.frame(2160, 3840, 2).reshape(-1, 3)
    init = img[np.random.RandomState(0).choice(len(img), 16, replace=False)].astype(np.float64)
    lo, hi = D.shard_pixels(len(img), rank, world)
    comm = D.nccl_comm()
    shard = _capi.DeviceBuffer((hi - lo) * 3).upload(np.ascontiguousarray(img[lo:hi]))
    iters = int(os.environ.get("DP_CHECK_ITERS", "20"))

    def timed(kw):
        kmeans.lloyd_device(shard.ptr, hi - lo, init, -1.0, 2, **kw())      # warm-up
        best, out = None, None
        for rep in range(3):
            args = kw()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = kmeans.lloyd_device(shard.ptr, hi - lo, init, -1.0, iters, check_every=iters, **args)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        t = torch.tensor([best], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    dt_nccl, res_nccl = timed(kw=lambda: {"comm": comm})
    dt, res = dt_nccl, res_nccl
    if world > 1:
        dt, res = timed(kw=lambda: {"p2p": D.p2p_exchange()})
        if rank == 0:
            print(f"[kmeans] world={world}: NCCL exchange {dt_nccl/iters*1e6:.1f} us/iteration, "
                  f"peer-memory exchange {dt/iters*1e6:.1f} us/iteration; identical centres: "
                  f"{bool(np.array_equal(res[0], res_nccl[0]))}")
    cent = res[0]
    ok_km = True
    if rank == 0:
        buf = _capi.DeviceBuffer(img.nbytes).upload(np.ascontiguousarray(img))
        t0 = time.perf_counter()
        ref = kmeans.lloyd_device(buf.ptr, len(img), init, -1.0, iters, check_every=iters)
        t1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = kmeans.lloyd_device(buf.ptr, len(img), init, -1.0, iters, check_every=iters)
        t1 = min(t1, time.perf_counter() - t0)
        ok_km = bool(np.array_equal(ref[0], cent)) and ref.ties == res.ties
        print(f"[kmeans] world={world}: {iters} Lloyd iterations over 8.29 Mpx sharded: {dt*1e3:.3f} ms "
              f"({dt/iters*1e6:.1f} us/iteration, max over ranks); 1-GPU full image: {t1*1e3:.3f} ms "
              f"({t1/iters*1e6:.1f} us/iteration); centres identical to the 1-GPU run: {ok_km}; ties {res.ties}")
    # ---- frame-sharded dithering --------------------------------------------------------------
    pal = synth.hex_palette(synth.PICO8)
    frames = np.stack([synth.frame(270, 480, 100 + t) for t in range(16)])
    out = D.process_frames_sharded(frames, lambda a: engine.dither_frames(a, pal, "bayer", {"size": "8x8"})
                                   if len(a) else a)
    ok_fr = True
    if rank == 0:
        ref = engine.dither_frames(frames, pal, "bayer", {"size": "8x8"})
        ok_fr = bool(np.array_equal(ref, out))
        print(f"[frames] world={world} 16 frames sharded, identical to the 1-GPU run: {ok_fr}")
    flag = torch.tensor([int(ok_km and ok_fr)], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)


if __name__ == "__main__":
    main()
