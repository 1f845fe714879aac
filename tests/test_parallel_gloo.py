"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: sharding, the broadcast rotation-dropout draw and
the gradient all-reduce.  The STN path itself exchanges nothing; shard invariance of its results is covered by
tests/test_oracle.py::test_shard_invariance (CPU) and tests/test_gpu_parity.py (device)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loans_b200 import parallel as P


def test_shard_bounds_cover_and_partition():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_keeps_crops_with_their_frames():
    x = torch.arange(6).view(6, 1, 1, 1)
    theta = torch.arange(6 * 4).view(24, 1, 1).expand(24, 2, 3)
    xs, ts = P.shard_batch(x, theta, 3, 1, crops_per_frame=4)
    assert xs.flatten().tolist() == [2, 3] and ts[:, 0, 0].tolist() == list(range(8, 16))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank draws its own flag; all must end up with rank 0's
        mine = float(rank == 0)
        got = P.broadcast_mask_value(mine, src=0)
        shapes = [(3, 4), (5,), (2, 2, 2)]
        grads = [torch.full(s, float(rank + 1)) * (i + 1) for i, s in enumerate(shapes)]
        ar = P.GradientAllReduce(shapes, "cpu")
        views = ar.start(grads).finish()
        want = [(i + 1) * (1 + 2) / 2.0 for i in range(3)]
        ok = got == 1.0 and all(torch.allclose(v, torch.full_like(v, w)) for v, w in zip(views, want))
        lo, hi = P.shard_bounds(10, world, rank)
        ret[rank] = (ok, lo, hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_and_broadcast():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] == (True, 0, 5) and ret[1] == (True, 5, 10)


def test_single_process_is_a_no_op():
    ar = P.GradientAllReduce([(4,)], "cpu")
    v = ar.start([torch.arange(4.0)]).finish()
    assert np.allclose(v[0].numpy(), [0, 1, 2, 3])
    assert P.broadcast_mask_value(0.0) == 0.0
