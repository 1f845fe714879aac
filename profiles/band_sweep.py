"""A/B timing of the backward kernels on a B200: band kernel knobs vs the general kernel, CUDA-graph replay over
rotating buffer sets (> L2), CUDA events.  Usage: python profiles/band_sweep.py [cfg2 cfg5 cfg3]  -> JSON lines."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402


def time_bwd(wl, sets, reps):
    L = _lib.lib()
    B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
    N = B * K
    bf16 = wl.out_dtype == "bf16"
    dt = _lib.BF16 if bf16 else _lib.F32

    def bwd(e):
        _lib.check(L.loans_stn_crop_bwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["gy"].data_ptr(), None,
                                        e["gtheta"].data_ptr(), e["gx"].data_ptr(), None, N, K, C, H, Wd, oH, oW, dt,
                                        torch.cuda.current_stream().cuda_stream), "crop_bwd")
    bwd(sets[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for e in sets:
            bwd(e)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(sets))


def main():
    names = sys.argv[1:] or ["cfg2", "cfg5", "cfg3"]
    dev = torch.device("cuda", 0)
    for name in names:
        wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
        if os.environ.get("BATCH"):
            wl = wl._replace(batch=int(os.environ["BATCH"]))
        B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
        N = B * K
        ydt = torch.bfloat16 if wl.out_dtype == "bf16" else torch.float32
        set_bytes = 8 * B * C * H * Wd + N * C * oH * oW * 4
        S = int(min(12, max(4, math.ceil(3.0 * 126e6 / set_bytes))))
        sets = []
        for s in range(S):
            d = W.make_inputs(wl, seed=77 + s)
            sets.append({"x": torch.from_numpy(d["x"]).to(dev), "theta": torch.from_numpy(d["theta"]).to(dev),
                         "gy": torch.from_numpy(d["gy"]).to(dev).to(ydt),
                         "gtheta": torch.empty((N, 2, 3), dtype=torch.float32, device=dev),
                         "gx": torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)})
        reps = 40 if wl.batch * wl.height <= 64 * 224 else 8
        _, bwd_bytes = W.algorithmic_bytes(wl, need_gx=True)
        _lib.band_backward(False)
        for tpw in [int(v) for v in os.environ.get("SWEEP_TPW", "0").split(",")]:
            _lib.check(_lib.lib().loans_stn_configure(10, tpw), "cfg")
            us = time_bwd(wl, sets, reps)
            print(json.dumps({"wl": name, "batch": wl.batch, "kernel": "general", "tiles_per_warp": tpw, "us": round(us, 2), "gbs": round(bwd_bytes / us / 1e3, 1)}), flush=True)
        _lib.check(_lib.lib().loans_stn_configure(10, 0), "cfg")
        _lib.band_backward(True)
        if os.environ.get("SWEEP", "x") == "":
            continue
        combos = os.environ.get("SWEEP", "1,3;0;0;0;0").split(";")
        variants, css, tiles, rowss, flagss = [[int(v) for v in c.split(",")] for c in combos]
        for variant0 in variants:
          for flags in flagss:
            variant = variant0 + 16 * flags
            for cs in css:
                for tile_kb in tiles:
                    for rows in rowss:
                        _lib.band_tuning(cs=cs, rows=rows, tile_kb=tile_kb, variant=variant)
                        us = time_bwd(wl, sets, reps)
                        print(json.dumps({"wl": name, "batch": wl.batch, "kernel": "band", "variant": variant, "cs": cs, "tile_kb": tile_kb,
                                          "rows": rows, "us": round(us, 2), "gbs": round(bwd_bytes / us / 1e3, 1)}), flush=True)
        _lib.band_tuning()
        _lib.band_backward(None)
        del sets
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
