"""What a WRONG 'upright' hint costs: rotated theta (r ~ U(-0.2, 0.2)) and upright theta through loans_stn_crop_bwd_ex with
mask01 = 1, with and without LOANS_STN_FLAG_UPRIGHT (the kernels for axis-aligned crops test every crop on the device and run
the general roles for rotated ones inside the same launch).  usage: hint_probe.py [cfg2 cfg5 cfg3 cfg4]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

L = _lib.lib()
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
L.loans_stn_crop_bwd_ex.argtypes = [vp, vp, cf, vp, vp, vp, vp, vp, vp] + [ci] * 9 + [vp]
dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["cfg2", "cfg5"]:
    wl = W.WORKLOADS[name]
    B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
    N = B * K
    bf16 = wl.out_dtype == "bf16"
    S = 2 if B * H * Wd > 3e7 else 6
    for rotate in (False, True):
        sets = []
        for s in range(S):
            d = W.make_inputs(wl._replace(rotation_ratio=None if rotate else 0.0), seed=31 + s, rotate=rotate)
            e = {k: torch.from_numpy(d[k]).to(dev) for k in ("x", "theta", "gy")}
            if bf16:
                e["gy"] = e["gy"].to(torch.bfloat16)
            e["gt"] = torch.empty((N, 2, 3), device=dev)
            e["gx"] = torch.empty((B, C, H, Wd), device=dev)
            sets.append(e)
        for flags in (0, 2):
            def bwd(e):
                assert L.loans_stn_crop_bwd_ex(e["x"].data_ptr(), e["theta"].data_ptr(), 1.0, e["gy"].data_ptr(), None, None, e["gt"].data_ptr(),
                                               e["gx"].data_ptr(), None, flags, N, K, C, H, Wd, oH, oW, 1 if bf16 else 0,
                                               torch.cuda.current_stream().cuda_stream) == 0
            bwd(sets[0])
            torch.cuda.synchronize()
            kern = _lib.last_kernel()
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
            print(json.dumps({"wl": name, "theta": "rotated" if rotate else "upright", "hint": bool(flags), "kernel": kern,
                              "bwd_us": round(e0.elapsed_time(e1) * 1e3 / (10 * S), 1)}), flush=True)
