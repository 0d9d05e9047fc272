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


def kmeans_fit_sharded(pixels_u8: np.ndarray, init_centers: np.ndarray, tol: float,
                       max_iter: int = kmeans.MAX_ITER) -> Tuple[np.ndarray, int]:
    """Full-image Lloyd iterations with the pixels sharded over the ranks of the job
    (BASELINE config 3, throughput mode).  Every rank passes the SAME ``pixels_u8`` [N,3] and
    initial centres; each uploads only its shard, accumulates exact integer sums on its GPU,
    all-reduces them (NCCL) and updates identical centres.  Returns (centres f64 [K,3], iters)."""
    import torch
    from . import _capi
    rank, world = init_process_group()
    pix = np.ascontiguousarray(pixels_u8, np.uint8).reshape(-1, 3)
    lo, hi = shard_pixels(pix.shape[0], rank, world)
    K = int(init_centers.shape[0])
    dev = torch.device("cuda", torch.cuda.current_device())
    shard = torch.from_numpy(pix[lo:hi]).to(dev)
    sums = torch.zeros(K * 4, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def reduce():
        allreduce_sums_(sums)

    import ctypes as C
    return kmeans.lloyd_device(shard.data_ptr(), hi - lo, init_centers, tol, max_iter,
                               sums_ptr=sums.data_ptr(), allreduce=reduce,
                               stream=C.c_void_p(stream.cuda_stream))


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
