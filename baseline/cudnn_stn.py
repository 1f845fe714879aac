"""The reference's GPU path for the STN stage, for comparison only (SURVEY.md section 8d "GPU comparison point").

On a GPU, chainer 4.1.0's F.spatial_transformer_grid / F.spatial_transformer_sampler call cuDNN's
cudnnSpatialTfGridGeneratorForward/Backward and cudnnSpatialTfSamplerForward/Backward (plus two layout copies: cuDNN
wants the grid as (N, oH, oW, 2), chainer hands out (N, 2, oH, oW)).  This module drives exactly those four cuDNN
entry points through ctypes on torch-owned buffers, so bench.py can time them on the same synthetic inputs next to
our kernels.  It is a measurement arm: nothing in loans_b200 imports it, and no product path calls cuDNN.
rotation_dropout (a scalar mask on theta) is applied with one torch multiply, as chainer would launch an elementwise
kernel for it.
"""
import ctypes

import torch

CUDNN_TENSOR_NCHW = 0
CUDNN_DATA_FLOAT = 0
CUDNN_SAMPLER_BILINEAR = 0


class CudnnStn:
    def __init__(self, b, c, h, w, oh, ow, device):
        self.lib = None
        for name in ("libcudnn.so.9", "libcudnn.so"):
            try:
                self.lib = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if self.lib is None:
            raise RuntimeError("libcudnn not found")
        L = self.lib
        self.shape = (b, c, h, w, oh, ow)
        self.device = device
        vp = ctypes.c_void_p
        self.handle = vp()
        self._ok(L.cudnnCreate(ctypes.byref(self.handle)), "cudnnCreate")
        self.st = vp()
        self._ok(L.cudnnCreateSpatialTransformerDescriptor(ctypes.byref(self.st)), "CreateSpatialTransformerDescriptor")
        dims = (ctypes.c_int * 4)(b, c, oh, ow)
        self._ok(L.cudnnSetSpatialTransformerNdDescriptor(self.st, CUDNN_SAMPLER_BILINEAR, CUDNN_DATA_FLOAT, 4, dims),
                 "SetSpatialTransformerNdDescriptor")
        self.xd, self.yd = vp(), vp()
        for d, (hh, ww) in ((self.xd, (h, w)), (self.yd, (oh, ow))):
            self._ok(L.cudnnCreateTensorDescriptor(ctypes.byref(d)), "CreateTensorDescriptor")
            self._ok(L.cudnnSetTensor4dDescriptor(d, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, b, c, hh, ww), "SetTensor4dDescriptor")
        self.one = ctypes.c_float(1.0)
        self.zero = ctypes.c_float(0.0)
        self.version = int(L.cudnnGetVersion())

    def _ok(self, status, what):
        if status != 0:
            raise RuntimeError("cuDNN %s failed with status %d" % (what, status))

    def set_stream(self, stream):
        self._ok(self.lib.cudnnSetStream(self.handle, ctypes.c_void_p(stream)), "cudnnSetStream")

    def forward(self, x, theta, mask01, y, grid_nhw2, grid_chainer):
        """theta (B,2,3) -> masked theta -> grid (B,oH,oW,2) -> y; grid_chainer (B,2,oH,oW) is the transpose chainer returns."""
        L, vp = self.lib, ctypes.c_void_p
        th = theta
        if mask01 != 1.0:
            th = theta.clone()
            th[:, 0, 1] *= mask01
            th[:, 1, 0] *= mask01
        self._ok(L.cudnnSpatialTfGridGeneratorForward(self.handle, self.st, vp(th.data_ptr()), vp(grid_nhw2.data_ptr())), "GridGeneratorForward")
        self._ok(L.cudnnSpatialTfSamplerForward(self.handle, self.st, ctypes.byref(self.one), self.xd, vp(x.data_ptr()),
                                                vp(grid_nhw2.data_ptr()), ctypes.byref(self.zero), self.yd, vp(y.data_ptr())),
                 "SamplerForward")
        grid_chainer.copy_(grid_nhw2.permute(0, 3, 1, 2))            # the layout copy chainer makes for its (B,2,oH,oW) output
        return th

    def backward(self, x, masked_theta_unused, mask01, gy, grid_nhw2, gx, dgrid_nhw2, gtheta):
        L, vp = self.lib, ctypes.c_void_p
        self._ok(L.cudnnSpatialTfSamplerBackward(self.handle, self.st, ctypes.byref(self.one), self.xd, vp(x.data_ptr()),
                                                 ctypes.byref(self.zero), self.xd, vp(gx.data_ptr()), ctypes.byref(self.one),
                                                 self.yd, vp(gy.data_ptr()), vp(grid_nhw2.data_ptr()), ctypes.byref(self.zero),
                                                 vp(dgrid_nhw2.data_ptr())), "SamplerBackward")
        self._ok(L.cudnnSpatialTfGridGeneratorBackward(self.handle, self.st, vp(dgrid_nhw2.data_ptr()), vp(gtheta.data_ptr())),
                 "GridGeneratorBackward")
        if mask01 != 1.0:
            gtheta[:, 0, 1] *= mask01
            gtheta[:, 1, 0] *= mask01


def time_cudnn(wl, sets, mask01, reps, device):
    """Median-free simple timing: CUDA-graph replay over the rotating sets (same hygiene as our arm).  K must be 1."""
    b, c, h, w, oh, ow = wl.batch, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
    stn = CudnnStn(b, c, h, w, oh, ow, device)
    bufs = []
    for e in sets:
        bufs.append({"grid2": torch.empty((b, oh, ow, 2), dtype=torch.float32, device=device),
                     "dgrid2": torch.empty((b, oh, ow, 2), dtype=torch.float32, device=device),
                     "y": torch.empty((b, c, oh, ow), dtype=torch.float32, device=device),
                     "gy": e["gy"].float().contiguous()})

    def step(e, t):
        stn.set_stream(torch.cuda.current_stream().cuda_stream)
        stn.forward(e["x"], e["theta"], mask01, t["y"], t["grid2"], e["grid"])
        stn.backward(e["x"], None, mask01, t["gy"], t["grid2"], e["gx"], t["dgrid2"], e["gtheta"])

    step(sets[0], bufs[0])
    torch.cuda.synchronize()
    y_ref, gx_ref, gt_ref = bufs[0]["y"].clone(), sets[0]["gx"].clone(), sets[0]["gtheta"].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(sets[0], bufs[0])                 # warm every cuDNN kernel on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for e, t in zip(sets, bufs):
            step(e, t)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(sets))
    return us, stn.version, (y_ref, gx_ref, gt_ref)
