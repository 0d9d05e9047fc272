"""N > 1 host logic on CPU: two processes, gloo backend (no GPU, no CUDA call).

Covers what the multi-GPU path adds on top of the single-GPU kernels: the rendezvous, the
contiguous frame sharding + ordered gather, and the integer all-reduce of k-means centroid sums
(shard-count invariance).  The per-shard arithmetic is done with numpy stand-ins here; the GPU
versions of the same steps are covered by tests/test_gpu_parity.py.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dither_pie_b200 import distributed as D
    r, w = D.init_process_group("gloo")
    assert (r, w) == (rank, world)

    # --- k-means: exact integer sums, all-reduced -> identical centres on every rank ----------
    rs = np.random.RandomState(3)
    pix = rs.randint(0, 256, size=(5003, 3)).astype(np.uint8)     # same on every rank
    cent = pix[rs.choice(len(pix), 8, replace=False)].astype(np.float64)
    lo, hi = D.shard_pixels(len(pix), rank, world)
    mine = pix[lo:hi].astype(np.int64)
    lab = ((mine[:, None, :] - cent[None]) ** 2).sum(2).argmin(1)
    sums = np.zeros((8, 4), np.int64)
    np.add.at(sums[:, :3], lab, mine)
    np.add.at(sums[:, 3], lab, 1)
    t = torch.from_numpy(sums.reshape(-1).copy())
    D.allreduce_sums_(t)
    full = pix.astype(np.int64)
    lab_all = ((full[:, None, :] - cent[None]) ** 2).sum(2).argmin(1)
    ref = np.zeros((8, 4), np.int64)
    np.add.at(ref[:, :3], lab_all, full)
    np.add.at(ref[:, 3], lab_all, 1)
    ok_sums = bool(np.array_equal(t.numpy().reshape(8, 4), ref))

    # --- frames: contiguous shards, ordered gather on rank 0 ---------------------------------
    frames = np.arange(7 * 4 * 5 * 3, dtype=np.uint8).reshape(7, 4, 5, 3)
    out = D.process_frames_sharded(frames, lambda a: 255 - a)
    ok_frames = True
    if rank == 0:
        ok_frames = bool(np.array_equal(out, 255 - frames))
    else:
        ok_frames = out is None
    q.put((rank, ok_sums, ok_frames))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gloo_sharding_and_integer_allreduce():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
