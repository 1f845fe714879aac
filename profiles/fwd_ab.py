"""A/B of the forward's pixels-per-CTA knob (key 12) on one workload: forward alone and the fwd+bwd step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch  # noqa: E402

import so_ab as SA  # noqa: E402
from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
N = B * K
dev = torch.device("cuda", 0)
L = _lib.lib()
sets = []
for s in range(6 if B <= 64 else 4):
    d = W.make_inputs(wl, seed=77 + s)
    sets.append({"x": torch.from_numpy(d["x"]).to(dev), "theta": torch.from_numpy(d["theta"]).to(dev),
                 "gy": torch.from_numpy(d["gy"]).to(dev), "y": torch.empty((N, C, oH, oW), device=dev),
                 "grid": torch.empty((N, 2, oH, oW), device=dev), "gtheta": torch.empty((N, 2, 3), device=dev),
                 "gx": torch.empty((B, C, H, Wd), device=dev)})
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def fwd(e):
    assert L.loans_stn_crop_fwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["y"].data_ptr(), e["grid"].data_ptr(), N, K, C, H, Wd, oH, oW, 0, st()) == 0


def bwd(e):
    assert L.loans_stn_crop_bwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["gy"].data_ptr(), None, e["gtheta"].data_ptr(),
                                e["gx"].data_ptr(), None, N, K, C, H, Wd, oH, oW, 0, st()) == 0


reps = 40 if B <= 64 else 8
for px in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,256,512,768,1024,2048,0").split(",")]:
    _lib.check(L.loans_stn_configure(12, px), "cfg")
    print(json.dumps({"wl": name, "fwd_px_per_cta": px, "fwd_us": round(SA.graph_time(fwd, sets, reps), 2),
                      "step_us": round(SA.graph_time(lambda e: (fwd(e), bwd(e)), sets, reps), 2)}), flush=True)
