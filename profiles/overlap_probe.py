"""Can the theta kernel (read-bound) and a gx-only row-band kernel (write-bound) overlap?  Uses an A/B build (profiles/ab_local.sh gxonly "-DSTN_BAND_GXONLY") with
-DSTN_BAND_GXONLY (ab_builds/gxonly.so): crop_bwd with gx runs the gx-only band kernel, crop_bwd without gx the table-driven theta
kernel; timed back to back on one stream and forked onto two streams (CUDA-graph replay).  usage: overlap_probe.py [cfg5 ...]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import workloads as W  # noqa: E402

vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
L = ctypes.CDLL(os.path.join(ROOT, "ab_builds", "gxonly.so"))
L.loans_stn_crop_bwd.argtypes = [vp, vp, cf, vp, vp, vp, vp, vp] + [ci] * 8 + [vp]
dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["cfg5", "cfg2"]:
    wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
    B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
    N = B * K
    S = 2 if B >= 512 else 6
    sets = []
    for s in range(S):
        d = W.make_inputs(wl, seed=7 + s)
        sets.append({k: torch.from_numpy(d[k]).to(dev) for k in ("x", "theta", "gy")})
        sets[-1]["gt"] = torch.empty((N, 2, 3), device=dev)
        sets[-1]["gx"] = torch.empty((B, C, H, Wd), device=dev)
    side = torch.cuda.Stream()

    def call(e, gx, stream):
        assert L.loans_stn_crop_bwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["gy"].data_ptr(), None, e["gt"].data_ptr(),
                                    e["gx"].data_ptr() if gx else None, None, N, K, C, H, Wd, oH, oW, 0, stream.cuda_stream) == 0

    def serial(e):
        cur = torch.cuda.current_stream()
        call(e, False, cur)
        call(e, True, cur)

    def forked(e):
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        call(e, False, side)
        call(e, True, cur)
        cur.wait_stream(side)

    def only(gx):
        return lambda e: call(e, gx, torch.cuda.current_stream())

    for label, fn in (("theta kernel", only(False)), ("gx-only band kernel", only(True)), ("back to back", serial), ("two streams", forked)):
        fn(sets[0])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for e in sets:
                fn(e)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"wl": name, "what": label, "us": round(e0.elapsed_time(e1) * 1e3 / (10 * S), 1)}), flush=True)
