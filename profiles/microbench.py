"""Calibration on the GPU box: per-kernel floor inside a CUDA graph, and plain fill / copy speed at the gx size.
Same harness style as bench.py (graphs of 5 launches on rotating buffers, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loans_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
S = 5
theta = [torch.randn(64, 2, 3, device=dev) for _ in range(S)]
out = [torch.empty(64, 2, 3, device=dev) for _ in range(S)]
gx = [torch.empty(64, 3, 224, 224, device=dev) for _ in range(S)]
src = [torch.randn(64, 3, 224, 224, device=dev) for _ in range(S)]


def timeit(fn, name, reps=200):
    g = torch.cuda.CUDAGraph()
    fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        fn()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %.2f us per launch" % (name, e0.elapsed_time(e1) * 1e3 / (reps * S)))


def tiny():
    for i in range(S):
        L.loans_stn_rotation_dropout(theta[i].data_ptr(), 0.5, out[i].data_ptr(), 64, torch.cuda.current_stream().cuda_stream)


timeit(tiny, "tiny kernel (2 CTAs, 384 floats)")
timeit(lambda: [g.zero_() for g in gx], "torch zero_ of 38.5 MB (rotating 5 buffers)")
timeit(lambda: [g.copy_(s) for g, s in zip(gx, src)], "torch copy_ 38.5 MB -> 38.5 MB")
big = [torch.empty(256, 3, 512, 512, device=dev) for _ in range(2)]
timeit(lambda: [g.zero_() for g in big], "torch zero_ of 805 MB", reps=20)
