"""A/B of two builds of libloans_stn.so on the same GPU, interleaved: times loans_stn_crop_fwd / _bwd (CUDA-graph replay over
rotating buffer sets).  usage: so_ab.py <a.so> <b.so> [<c.so> ...] [cfg2 cfg5 ...]"""
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


def load(path):
    L = ctypes.CDLL(path)
    L.loans_stn_crop_fwd.argtypes = [vp, vp, cf, vp, vp] + [ci] * 8 + [vp]
    L.loans_stn_crop_bwd.argtypes = [vp, vp, cf, vp, vp, vp, vp, vp] + [ci] * 8 + [vp]
    return L


def graph_time(fn, sets, reps):
    fn(sets[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for e in sets:
            fn(e)
    for _ in range(3):
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
    libs = [(p, load(p)) for p in sys.argv[1:] if p.endswith(".so")]
    names = [a for a in sys.argv[1:] if not a.endswith(".so")] or ["cfg2", "cfg5"]
    dev = torch.device("cuda", 0)
    for name in names:
        wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
        B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
        N = B * K
        bf16 = wl.out_dtype == "bf16"
        ydt = torch.bfloat16 if bf16 else torch.float32
        dt = 1 if bf16 else 0
        S = int(min(12, max(4, math.ceil(3.0 * 126e6 / (8 * B * C * H * Wd + N * C * oH * oW * 4)))))
        sets = []
        for s in range(S):
            d = W.make_inputs(wl, seed=77 + s)
            sets.append({"x": torch.from_numpy(d["x"]).to(dev), "theta": torch.from_numpy(d["theta"]).to(dev),
                         "gy": torch.from_numpy(d["gy"]).to(dev).to(ydt), "y": torch.empty((N, C, oH, oW), dtype=ydt, device=dev),
                         "grid": torch.empty((N, 2, oH, oW), dtype=torch.float32, device=dev),
                         "gtheta": torch.empty((N, 2, 3), dtype=torch.float32, device=dev),
                         "gx": torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev)})
        reps = 40 if B <= 64 else 8
        for rnd in range(2):
            for path, L in libs:
                st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731

                def fwd(e, L=L):
                    assert L.loans_stn_crop_fwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["y"].data_ptr(), e["grid"].data_ptr(),
                                                N, K, C, H, Wd, oH, oW, dt, st()) == 0

                def bwd(e, L=L):
                    assert L.loans_stn_crop_bwd(e["x"].data_ptr(), e["theta"].data_ptr(), 0.0, e["gy"].data_ptr(), None,
                                                e["gtheta"].data_ptr(), e["gx"].data_ptr(), None, N, K, C, H, Wd, oH, oW, dt, st()) == 0
                print(json.dumps({"wl": name, "so": os.path.relpath(path, ROOT), "fwd_us": round(graph_time(fwd, sets, reps), 2),
                                  "bwd_us": round(graph_time(bwd, sets, reps), 2),
                                  "step_us": round(graph_time(lambda e: (fwd(e), bwd(e)), sets, reps), 2)}), flush=True)
        del sets
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
