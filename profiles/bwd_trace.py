"""Debug: per-CTA timestamps of the general backward (library built with EXTRA=-DSTN_BAND_TRACE).  Runs fwd then bwd
back to back, as the training step does, after an L2 flush."""
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
wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
N = B * K
dev = torch.device("cuda", 0)
d = W.make_inputs(wl, seed=77)
x, th = torch.from_numpy(d["x"]).to(dev), torch.from_numpy(d["theta"]).to(dev)
gy = torch.from_numpy(d["gy"]).to(dev)
y = torch.empty((N, C, oH, oW), dtype=torch.float32, device=dev)
grid = torch.empty((N, 2, oH, oW), dtype=torch.float32, device=dev)
gt = torch.empty((N, 2, 3), dtype=torch.float32, device=dev)
gx = torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
_lib.band_backward(False)
L = _lib.lib()
trace = torch.zeros((8192, 8), dtype=torch.int64, device=dev)
L.loans_stn_debug_bwd_trace.argtypes = [ctypes.c_void_p]
assert L.loans_stn_debug_bwd_trace(trace.data_ptr()) == 0
st = torch.cuda.current_stream().cuda_stream
for it in range(3):
    flush.zero_()
    trace.zero_()
    _lib.check(L.loans_stn_crop_fwd(x.data_ptr(), th.data_ptr(), 0.0, y.data_ptr(), grid.data_ptr(), N, K, C, H, Wd, oH, oW, 0, st), "f")
    _lib.check(L.loans_stn_crop_bwd(x.data_ptr(), th.data_ptr(), 0.0, gy.data_ptr(), None, gt.data_ptr(), gx.data_ptr(), None,
                                    N, K, C, H, Wd, oH, oW, 0, st), "crop_bwd")
    torch.cuda.synchronize()
t = trace.cpu().numpy()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print("CTAs", t.shape[0], "kernel span %.2f us" % ((t[:, :3].max() - t0) / 1e3))
for role, nm in ((1, "gx"), (2, "theta")):
    m = t[:, 4] == role
    if not m.any():
        continue
    r = t[m]
    for k, lab in enumerate(("entry", "prologue", "role done")):
        col = (r[:, k] - t0) / 1e3
        print("%-6s %-10s min %6.2f median %6.2f p90 %6.2f max %6.2f" % (nm, lab, col.min(), np.median(col), np.percentile(col, 90), col.max()))
    dur = (r[:, 2] - r[:, 0]) / 1e3
    print("%-6s CTA duration median %.2f max %.2f; n=%d" % (nm, np.median(dur), dur.max(), m.sum()))
hist, edges = np.histogram((t[:, 0] - t0) / 1e3, bins=12)
print("entry histogram (us):", [(round(float(e), 1), int(h)) for e, h in zip(edges, hist)])
hist, edges = np.histogram((t[:, 2] - t0) / 1e3, bins=12)
print("exit histogram (us):", [(round(float(e), 1), int(h)) for e, h in zip(edges, hist)])
