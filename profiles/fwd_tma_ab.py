"""A/B of the forward kernels in a -DSTN_DEVEL build: direct gather vs the TMA-staged forward of axis-aligned crops
(LOANS_STN_CFG_TMA_FORWARD), CUDA-graph replay over rotating buffer sets.  usage: fwd_tma_ab.py ab_builds/devel.so [cfg3 cfg2 ...]"""
import ctypes
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import workloads as W  # noqa: E402

vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
L = ctypes.CDLL(sys.argv[1])
L.loans_stn_crop_fwd.argtypes = [vp, vp, cf, vp, vp] + [ci] * 8 + [vp]
dev = torch.device("cuda", 0)
for name in sys.argv[2:] or ["cfg3", "cfg2", "cfg5"]:
    wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
    B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
    N = B * K
    bf16 = wl.out_dtype == "bf16"
    ydt = torch.bfloat16 if bf16 else torch.float32
    S = int(min(8, max(2, math.ceil(3.0 * 126e6 / (4 * B * C * H * Wd)))))
    sets = []
    for s in range(S):
        sets.append({"x": torch.rand((B, C, H, Wd), device=dev), "theta": torch.from_numpy(W.make_theta(__import__("numpy").random.default_rng(s), N, rotate=False)).to(dev),
                     "y": torch.empty((N, C, oH, oW), dtype=ydt, device=dev), "grid": torch.empty((N, 2, oH, oW), device=dev)})
    for tma in (0, 1, 0, 1):
        assert L.loans_stn_configure(2, tma) == 0

        def fwd(e):
            assert L.loans_stn_crop_fwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["y"].data_ptr(), e["grid"].data_ptr(), N, K, C, H, Wd, oH, oW,
                                        1 if bf16 else 0, torch.cuda.current_stream().cuda_stream) == 0
        fwd(sets[0])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for e in sets:
                fwd(e)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"wl": name, "tma_forward": tma, "fwd_us": round(e0.elapsed_time(e1) * 1e3 / (10 * S), 2)}), flush=True)
    L.loans_stn_configure(2, 0)
