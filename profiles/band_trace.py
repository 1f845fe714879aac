"""Debug: per-CTA stage timestamps of the band kernel (library built with EXTRA=-DSTN_BAND_TRACE)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
a = [int(v) for v in sys.argv[2:]] + [0] * 5
wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
N = B * K
dev = torch.device("cuda", 0)
d = W.make_inputs(wl, seed=77)
x, th = torch.from_numpy(d["x"]).to(dev), torch.from_numpy(d["theta"]).to(dev)
gy = torch.from_numpy(d["gy"]).to(dev)
gt = torch.empty((N, 2, 3), dtype=torch.float32, device=dev)
gx = torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
_lib.band_tuning(variant=a[0], cs=a[1], rows=a[2], tile_kb=a[3])
_lib.band_backward(True)
cs = a[1] or 8
L = _lib.lib()
trace = torch.zeros((N * cs, 16), dtype=torch.int64, device=dev)
L.loans_stn_debug_band_trace.argtypes = [ctypes.c_void_p]
assert L.loans_stn_debug_band_trace(trace.data_ptr()) == 0
for it in range(3):
    flush.zero_()
    trace.zero_()
    _lib.check(L.loans_stn_crop_bwd(x.data_ptr(), th.data_ptr(), 0.0, gy.data_ptr(), None, gt.data_ptr(), gx.data_ptr(), None,
                                    N, K, C, H, Wd, oH, oW, _lib.F32, torch.cuda.current_stream().cuda_stream), "crop_bwd")
    torch.cuda.synchronize()
t = trace.cpu().numpy()
t0 = t[:, 0].min()
names = ["entry", "crop ok", "prologue", "zero spans", "tile zeroed", "loads back", "reduced", "scattered", "fenced+sync",
         "band done", "gtheta", "read-wait"]
print("CTAs", t.shape[0], "kernel span %.2f us" % ((t[:, :12].max() - t0) / 1e3))
for k, nm in enumerate(names):
    col = t[:, k]
    col = col[col > 0] - t0
    if len(col):
        print("%-12s min %7.2f  median %7.2f  p90 %7.2f  max %7.2f us" % (nm, col.min() / 1e3, np.median(col) / 1e3,
                                                                          np.percentile(col, 90) / 1e3, col.max() / 1e3))
dur = (t[:, 11] - t[:, 0]) / 1e3
print("CTA duration: median %.2f  max %.2f us" % (np.median(dur), dur.max()))
sm = t[:, 15]
print("distinct SMs", len(np.unique(sm)), "max CTAs on one SM", np.bincount(sm.astype(int)).max())
for k in range(1, 12):
    dlt = (t[:, k] - t[:, k - 1]) / 1e3
    print("stage %-12s median %6.2f  max %6.2f us" % (names[k], np.median(dlt), dlt.max()))
print("by rank: median end of [prologue, first loads issued, tile zeroed, loads back, band done, gtheta]")
r = np.arange(t.shape[0]) % cs
for k in range(cs):
    m = r == k
    print("rank %d:" % k, " ".join("%6.2f" % (np.median(t[m, c] - t0) / 1e3) for c in (2, 3, 4, 5, 9, 10)),
          "  max band done %.2f" % ((t[m, 9] - t0).max() / 1e3))
slow = np.argsort(-(t[:, 9] - t0))[:12]
print("slowest CTAs (cta, rank, sm): timeline")
for c in slow:
    print(c, c % cs, int(t[c, 15]), " ".join("%6.2f" % ((t[c, k] - t0) / 1e3) for k in range(12)))
smc = np.bincount(sm.astype(int), minlength=148)
print("CTAs per SM histogram:", np.bincount(smc))
