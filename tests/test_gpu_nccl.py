"""Two-rank NCCL tests (one process per GPU, spawned here): the step's one collective -- GradientAllReduce over NCCL on a
side stream -- the broadcast rotation-dropout draw, and shard invariance of the STN path itself (the concatenation of the
shards' results is the unsharded result, bit for bit, with no collective on the data path).  Skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from loans_b200 import parallel as P
    from loans_b200 import workloads as W
    from loans_b200.functions import stn_crop
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        got = P.broadcast_mask_value(float(rank == 0), src=0)
        # the localizer-sized bucket (12.59 M fp32) plus two small tensors, mean over the ranks, on the side stream
        shapes = [(12_592_902,), (3, 4), (5,)]
        grads = [torch.full(s, float(rank + 1), device=dev) * (i + 1) for i, s in enumerate(shapes)]
        ar = P.GradientAllReduce(shapes, dev)
        ar.start(grads)
        busy = torch.ones(1 << 20, device=dev).mul_(2.0)                # work on the main stream while the collective runs
        views = ar.finish()
        want = [(i + 1) * sum(range(1, world + 1)) / world for i in range(3)]
        ok_ar = all(bool(torch.all(v == w)) for v, w in zip(views, want)) and float(busy[0]) == 2.0
        # shard invariance: every rank crops its own frames; nothing is exchanged on the data path
        wl = W.WORKLOADS["cfg1"]
        d = W.make_inputs(wl, batch=6, seed=11)
        lo, hi = P.shard_bounds(6, world, rank)
        x = torch.from_numpy(d["x"][lo:hi]).to(dev).requires_grad_()
        th = torch.from_numpy(d["theta"][lo:hi]).to(dev).requires_grad_()
        rois, points = stn_crop(x, th, (75, 75), mask01=got * 0.0)
        rois.backward(torch.from_numpy(d["gy"][lo:hi]).to(dev))
        torch.cuda.synchronize()
        ret[rank] = (got, ok_ar, lo, hi, rois.detach().cpu().numpy(), th.grad.cpu().numpy(), x.grad.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_allreduce_broadcast_and_shard_invariance():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from loans_b200 import workloads as W
    from oracle import stn_c as oc
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret[r][0] == 1.0 and ret[r][1] for r in range(world))
    assert (ret[0][2], ret[0][3], ret[1][2], ret[1][3]) == (0, 3, 3, 6)
    d = W.make_inputs(W.WORKLOADS["cfg1"], batch=6, seed=11)
    y0, _ = oc.crop_forward(d["x"], d["theta"], (75, 75), 0.0)
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], (75, 75), d["gy"], None, 0.0)
    y = np.concatenate([ret[0][4], ret[1][4]])
    gt = np.concatenate([ret[0][5], ret[1][5]])
    gx = np.concatenate([ret[0][6], ret[1][6]])
    assert np.array_equal(y, y0)
    assert np.abs(gt - gt0).max() <= 1e-4 * np.abs(gt0).max() and np.abs(gx - gx0).max() <= 2e-6 * np.abs(gx0).max()
