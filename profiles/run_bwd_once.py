"""Launch the backward of one workload a few times (for ncu).  usage: run_bwd_once.py cfg2 [variant cs rows tile_kb band(0/1)]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
a = [int(v) for v in sys.argv[2:]] + [0] * 5
wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
N = B * K
dev = torch.device("cuda", 0)
ydt = torch.bfloat16 if wl.out_dtype == "bf16" else torch.float32
d = W.make_inputs(wl, seed=77)
x, th = torch.from_numpy(d["x"]).to(dev), torch.from_numpy(d["theta"]).to(dev)
gy = torch.from_numpy(d["gy"]).to(dev).to(ydt)
gt = torch.empty((N, 2, 3), dtype=torch.float32, device=dev)
gx = torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)
_lib.band_tuning(variant=a[0], cs=a[1], rows=a[2], tile_kb=a[3])
_lib.band_backward(None if len(sys.argv) <= 6 else a[4] != 0)          # automatic dispatch unless the fifth knob is given
for _ in range(4):
    _lib.check(_lib.lib().loans_stn_crop_bwd(x.data_ptr(), th.data_ptr(), 0.0, gy.data_ptr(), None, gt.data_ptr(), gx.data_ptr(), None,
                                             N, K, C, H, Wd, oH, oW, _lib.BF16 if ydt == torch.bfloat16 else _lib.F32,
                                             torch.cuda.current_stream().cuda_stream), "crop_bwd")
torch.cuda.synchronize()
