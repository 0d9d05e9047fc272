"""Multi-GPU plumbing: one process per GPU (torchrun), ``torch.distributed`` for the rendezvous.

The hot path shards by FRAME with no data-path collective (video_processor.py:304-346 maps frames
over a process pool; here the pool is the 8 B200s of one box).  The only collective in the design
is the all-reduce of the K x 4 integer centroid sums of a pixel-sharded k-means.

``torch`` is imported lazily: the single-GPU product path does not need it.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import numpy as np

from . import kmeans
from .video_processor import shard_frames


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: Optional[str] = None):
    """Join the job described by the environment.  NCCL when a GPU is present, else gloo.
    Returns (rank, world).  A no-op for a single process."""
    rank, world, local = env_rank_world()
    if world == 1:
        return rank, world
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def barrier() -> None:
    """Process-group barrier (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def allreduce_sums_(sums) -> None:
    """In-place SUM all-reduce of an int64 tensor of per-cluster (sum r, sum g, sum b, count).
    Integers: the result -- hence every centre -- is independent of the number of shards and of
    the reduction order.  Works on CUDA tensors (NCCL over NVLink) and CPU tensors (gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)


def shard_pixels(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous pixel range of ``rank`` (same rule as frames)."""
    return shard_frames(n, rank, world)


_nccl_comm = None


def nccl_comm():
    """This rank's communicator for the library's own NCCL calls (created once per process: rank
    0 draws the unique id, torch.distributed broadcasts the 128 bytes, every rank joins).  The
    library dlopens the libnccl the process already carries (PyTorch's)."""
    global _nccl_comm
    if _nccl_comm is not None:
        return _nccl_comm
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ._capi import check, lib
    rank, world = init_process_group()
    if world == 1:
        return None
    # prefer the very file PyTorch loaded (nvidia-nccl wheel; a namespace package without __file__)
    import glob
    import sys
    cands = []
    for base in sys.path:
        cands += glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
    check(lib().dp_nccl_load(cands[0].encode() if cands else None), "dp_nccl_load")
    ident = np.zeros(128, np.uint8)
    if rank == 0:
        check(lib().dp_nccl_unique_id(ident.ctypes.data), "dp_nccl_unique_id")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(ident).to(dev)
    dist.broadcast(t, src=0)
    ident = t.cpu().numpy()
    h = C.c_void_p()
    check(lib().dp_nccl_comm_create(ident.ctypes.data, rank, world, C.byref(h)), "dp_nccl_comm_create")
    _nccl_comm = h.value
    return _nccl_comm


_p2p = None


def p2p_exchange():
    """(rank, world, inbox pointers, epoch) for ``kmeans.lloyd_device(p2p=...)``: every rank
    allocates an inbox in its own device memory, the ranks all-gather the 64-byte cudaIpc handles
    through ``torch.distributed`` and map each other's inboxes (peer access over NVLink).  Created
    once per process; each call returns a fresh epoch and passes a barrier, as dp_kmeans_lloyd_p2p
    requires between two uses of the same inboxes.  None for a single process."""
    global _p2p
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ._capi import check, lib
    rank, world = init_process_group()
    if world == 1:
        return None
    if world > 8:
        raise RuntimeError("the peer-memory exchange covers one NVSwitch domain (<= 8 ranks)")
    if _p2p is None:
        L = lib()
        mine = C.c_void_p()
        handle = np.zeros(64, np.uint8)
        check(L.dp_p2p_alloc(int(L.dp_p2p_inbox_bytes()), C.byref(mine), handle.ctypes.data), "dp_p2p_alloc")
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(handle).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(mine.value)
            else:
                h = np.ascontiguousarray(parts[r].cpu().numpy())
                q = C.c_void_p()
                check(L.dp_p2p_open(h.ctypes.data, C.byref(q)), "dp_p2p_open")
                ptrs.append(q.value)
        _p2p = {"rank": rank, "world": world, "ptrs": ptrs, "epoch": 0}
    _p2p["epoch"] += 1
    dist.barrier()
    if dist.get_backend() == "nccl":
        torch.cuda.synchronize()       # the barrier's own kernel has finished: every peer is past its last use
    return (_p2p["rank"], _p2p["world"], _p2p["ptrs"], _p2p["epoch"])


def kmeans_fit_sharded(pixels_u8: np.ndarray, init_centers: np.ndarray, tol: float,
                       max_iter: int = kmeans.MAX_ITER, check_every: int = 8, exchange: str = "p2p"):
    """Full-image Lloyd iterations with the pixels sharded over the ranks of the job
    (BASELINE config 3, throughput mode).  Every rank passes the SAME ``pixels_u8`` [N,3] and
    initial centres; each uploads only its shard; the Lloyd loop accumulates exact integer sums
    on its GPU and exchanges them per iteration without host synchronisation -- ``exchange="p2p"``:
    the kernels push the sums into the peers' inboxes over NVLink (dp_kmeans_lloyd_p2p);
    ``"nccl"``: ncclAllReduce on the kernel stream.  Returns (centres f64 [K,3], iters) --
    identical on all ranks."""
    from . import _capi
    rank, world = init_process_group()
    _capi.ensure_device()
    pix = np.ascontiguousarray(pixels_u8, np.uint8).reshape(-1, 3)
    lo, hi = shard_pixels(pix.shape[0], rank, world)
    buf = _capi.DeviceBuffer(max((hi - lo) * 3, 16)).upload(np.ascontiguousarray(pix[lo:hi]))
    try:
        if exchange == "p2p" and world <= 8:
            return kmeans.lloyd_device(buf.ptr, hi - lo, init_centers, tol, max_iter, p2p=p2p_exchange(),
                                       check_every=check_every)
        return kmeans.lloyd_device(buf.ptr, hi - lo, init_centers, tol, max_iter, comm=nccl_comm(),
                                   check_every=check_every)
    finally:
        buf.free()


def process_frames_sharded(frames: np.ndarray, run_shard: Callable[[np.ndarray], np.ndarray],
                           gather: bool = True) -> Optional[np.ndarray]:
    """Frame-sharded map: every rank holds the same ``frames`` [F,...]; rank r runs
    ``run_shard`` on its contiguous range.  With ``gather`` rank 0 returns the whole result in
    frame order (host-side gather through the process group), the other ranks return None."""
    rank, world = init_process_group()
    lo, hi = shard_frames(frames.shape[0], rank, world)
    out = run_shard(frames[lo:hi])
    if world == 1 or not gather:
        return out
    import torch.distributed as dist
    parts = [None] * world
    dist.gather_object(out, parts if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    return np.concatenate([p for p in parts if p is not None and len(p)], axis=0)
