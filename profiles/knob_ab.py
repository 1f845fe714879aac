"""A/B of the general backward's runtime knobs on one workload: (theta_first, tiles_per_warp) pairs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch  # noqa: E402

import band_sweep as BS  # noqa: E402
from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
N = B * K
dev = torch.device("cuda", 0)
ydt = torch.bfloat16 if wl.out_dtype == "bf16" else torch.float32
sets = []
for s in range(6):
    d = W.make_inputs(wl, seed=77 + s)
    sets.append({"x": torch.from_numpy(d["x"]).to(dev), "theta": torch.from_numpy(d["theta"]).to(dev),
                 "gy": torch.from_numpy(d["gy"]).to(dev).to(ydt), "gtheta": torch.empty((N, 2, 3), dtype=torch.float32, device=dev),
                 "gx": torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)})
_lib.band_backward(False)
for tf, tpw in ((0, 0), (1, 1), (1, 2), (1, 3), (0, 2), (0, 0)):
    _lib.theta_first(tf)
    _lib.check(_lib.lib().loans_stn_configure(10, tpw), "cfg")
    print(json.dumps({"wl": name, "theta_first": tf, "tiles_per_warp": tpw, "us": round(BS.time_bwd(wl, sets, 40), 2)}), flush=True)
