"""Data-parallel plumbing around the STN crop path: one process per GPU, frames sharded by batch index.

The path itself needs NO collective -- every crop depends only on its own frame and theta.
What a batch-sharded LoANs step needs around it is provided here:

* ``shard_bounds`` / ``shard_batch``      contiguous split of the frame batch (rank r takes [r*B/N, (r+1)*B/N));
* ``broadcast_mask_value``                rotation dropout draws ONE Bernoulli flag per call for the whole batch
                                          (reference functions/rotation_droput.py:41); with the batch spread over
                                          ranks the draw of rank 0 is broadcast so all shards see the same flag;
* ``GradientAllReduce``                   the one collective of the step: mean all-reduce of the localizer's
                                          gradients as a single flat fp32 bucket over NCCL (NVLink 5 / NVSwitch),
                                          launched on a side stream so it overlaps whatever the caller does next.
                                          (The reference never trained LoANs multi-GPU; its side project uses
                                          Chainer's MultiprocessParallelUpdater, schaaaafrichter/train.py:159-191.)

Works with the ``gloo`` backend on CPU tensors too (tests), ``nccl`` on the B200s.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_frames, world_size, rank):
    """[lo, hi) of the frames rank ``rank`` owns; sizes differ by at most one, earlier ranks get the extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(n_frames, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, theta, world_size, rank, crops_per_frame=1, *per_crop):
    """Slice frames ``x`` and the per-crop tensors (theta, plus any of gy / ggrid ...) for one rank."""
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    k = crops_per_frame
    out = [x[lo:hi], theta[lo * k:hi * k]]
    out.extend(t[lo * k:hi * k] for t in per_crop)
    return tuple(out)


def broadcast_mask_value(value, src=0, group=None, device=None):
    """Make every rank use rank ``src``'s rotation-dropout draw."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float32, device=device)
    dist.broadcast(t, src=src, group=group)
    return float(t.item())


class GradientAllReduce(object):
    """Mean all-reduce of a fixed set of gradient tensors through one flat fp32 bucket.

    ``start()`` packs the gradients and launches the all-reduce asynchronously (on its own CUDA stream when the
    tensors live on a GPU, after waiting for the producer stream); ``finish()`` waits, scales by 1/world and
    unpacks.  Sizes for LoANs: 12.59 M parameters (50 MB) for the ResNet-18 localizer at 224 px, 36.2 M (145 MB)
    at 512 px.
    """

    def __init__(self, shapes, device, group=None, dtype=torch.float32):
        self.group = group
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.flat = torch.zeros(sum(self.sizes), dtype=dtype, device=device)
        self.views = []
        off = 0
        for s, n in zip(self.shapes, self.sizes):
            self.views.append(self.flat[off:off + n].view(s))
            off += n
        self.is_cuda = self.flat.is_cuda
        self.stream = torch.cuda.Stream(device=device) if self.is_cuda else None
        self._work = None
        self._avg = False

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def start(self, grads=None):
        if grads is not None:
            for v, g in zip(self.views, grads):
                v.copy_(g)
        if self.world == 1:
            return self
        if self.is_cuda:
            # NCCL averages inside the collective (ReduceOp.AVG): no separate 1/N pass over the bucket
            self._avg = dist.get_backend(self.group) == "nccl"
            self.stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.stream):
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM,
                                             group=self.group, async_op=True)
        else:
            self._avg = False
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return self

    def finish(self, out=None):
        if self._work is not None:
            self._work.wait()
            self._work = None
            if self.is_cuda:
                torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
            if not self._avg:
                self.flat.mul_(1.0 / self.world)
        if out is not None:
            for v, g in zip(self.views, out):
                g.copy_(v)
        return self.views
