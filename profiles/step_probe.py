"""How much of the row-band backward's time at cfg5 (224 -> 75: crops step by 1.5 ... 2.7 frame pixels) goes to crops that step by
less than ~2 (P = 2 / Q = 2 phases, halo rows): the same batch with every scale drawn from [lo, hi].  usage: step_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

wl = W.WORKLOADS["cfg5"]._replace(rotation_ratio=0.0)
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
dev = torch.device("cuda", 0)
L = _lib.lib()
for lo, hi in ((0.5, 0.9), (0.72, 0.9), (0.5, 0.68), (0.8, 0.8)):
    sets = []
    for s in range(2):
        d = W.make_inputs(wl, seed=11 + s)
        rng = np.random.default_rng(5 + s)
        d["theta"][:, 0, 0] = rng.uniform(lo, hi, B).astype(np.float32)
        d["theta"][:, 1, 1] = rng.uniform(lo, hi, B).astype(np.float32)
        e = {k: torch.from_numpy(d[k]).to(dev) for k in ("x", "theta", "gy")}
        e["gt"] = torch.empty((B, 2, 3), device=dev)
        e["gx"] = torch.empty((B, C, H, Wd), device=dev)
        sets.append(e)

    def bwd(e):
        _lib.check(L.loans_stn_crop_bwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["gy"].data_ptr(), None, e["gt"].data_ptr(), e["gx"].data_ptr(),
                                        None, B, K, C, H, Wd, oH, oW, _lib.F32, torch.cuda.current_stream().cuda_stream), "bwd")
    bwd(sets[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for e in sets:
            bwd(e)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"scale": [lo, hi], "step_px": [round(lo * 223 / 74, 2), round(hi * 223 / 74, 2)], "kernel": _lib.last_kernel(),
                      "bwd_us": round(e0.elapsed_time(e1) * 1e3 / 20, 1)}), flush=True)
