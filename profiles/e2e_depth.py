"""e2e step time of HostCropPipeline at cfg2 against the number of device buffer sets (depth).  usage: e2e_depth.py [depths...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from loans_b200 import workloads as W  # noqa: E402
from loans_b200.pipeline import HostCropPipeline  # noqa: E402

wl = W.WORKLOADS["cfg2"]
d = W.make_inputs(wl, seed=1)
B, C, H, Wd, oH, oW = wl.batch, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
hx, hth, hgy = (torch.from_numpy(d[k]).pin_memory() for k in ("x", "theta", "gy"))
for depth in [int(a) for a in sys.argv[1:]] or [2, 3, 4, 2, 3]:
    pipe = HostCropPipeline(B, C, H, Wd, (oH, oW), need_gx=True, depth=depth)
    outs = [{"y": torch.empty((B, C, oH, oW)).pin_memory(), "grid": torch.empty((B, 2, oH, oW)).pin_memory(),
             "gtheta": torch.empty((B, 2, 3)).pin_memory(), "gx": torch.empty((B, C, H, Wd)).pin_memory()} for _ in range(depth)]
    for i in range(2 * depth):
        pipe.submit(hx, hth, hgy, outs[i % depth], mask01=0.0)
    pipe.drain()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pipe.s_in)
    for i in range(n):
        pipe.submit(hx, hth, hgy, outs[i % depth], mask01=0.0)
    e1.record(pipe.s_out)
    pipe.drain()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"depth": depth, "ms_per_step": round(ms, 4), "crops_per_s": round(B / ms * 1e3),
                      "h2d_gbs": round(pipe.h2d_bytes / ms / 1e6, 1), "d2h_gbs": round(pipe.d2h_bytes / ms / 1e6, 1)}), flush=True)
    del pipe
