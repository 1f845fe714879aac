"""Backward of BASELINE config 4 (16 crops per frame) through the C ABI, CUDA-graph replay over rotating sets: the general kernel,
the table-driven theta kernel alone (gx == NULL), and theta + kframe for several band heights, with and without programmatic
dependent launch.  usage: kframe_time.py [cfg4] [lib.so]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402

if len(sys.argv) > 2:
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
if os.environ.get("KF_BATCH"):
    wl = wl._replace(batch=int(os.environ["KF_BATCH"]))
B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
N = B * K
dev = torch.device("cuda", 0)
S = 2 if wl.batch >= 64 else 8
sets = []
for s in range(S):
    d = W.make_inputs(wl, seed=70 + s)
    sets.append({"x": torch.from_numpy(d["x"]).to(dev), "th": torch.from_numpy(d["theta"]).to(dev), "gy": torch.from_numpy(d["gy"]).to(dev).to(torch.bfloat16 if wl.out_dtype == "bf16" else torch.float32),
                 "gt": torch.empty((N, 2, 3), device=dev), "gx": torch.empty((B, C, H, Wd), device=dev)})
L = _lib.lib()


def bwd(e, gx=True):
    _lib.check(L.loans_stn_crop_bwd(e["x"].data_ptr(), e["th"].data_ptr(), 0.0, e["gy"].data_ptr(), None, e["gt"].data_ptr(),
                                    e["gx"].data_ptr() if gx else None, None, N, K, C, H, Wd, oH, oW, _lib.BF16 if wl.out_dtype == "bf16" else _lib.F32,
                                    torch.cuda.current_stream().cuda_stream), "crop_bwd")


def timed(fn, label):
    fn(sets[0])
    torch.cuda.synchronize()
    kern = _lib.last_kernel()
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
    print(json.dumps({"what": label, "kernel": kern, "us": round(e0.elapsed_time(e1) * 1e3 / (10 * S), 1)}), flush=True)


_lib.force_general(True)
timed(bwd, "general kernel")
_lib.force_general(False)
timed(lambda e: bwd(e, False), "theta only (no gx)")
for rows in [int(v) for v in os.environ.get("KF_ROWS", "0,16,32,64,128,256").split(",")] if K > 1 else []:
    _lib.kframe_rows(rows)
    timed(bwd, "theta + kframe, rows per CTA %d" % rows)
_lib.kframe_rows(0)
if K == 1:
    timed(bwd, "band backward (automatic rule)")
    L.loans_stn_configure(14, 1)
    for rows in [int(v) for v in os.environ.get("KF_ROWS", "0").split(",")]:
        _lib.kframe_rows(rows)
        timed(bwd, "theta + kframe on one crop per frame, rows per CTA %d" % rows)
    _lib.kframe_rows(0)
    L.loans_stn_configure(14, 0)
    sys.exit(0)
L.loans_stn_configure(1, 0)                      # LOANS_STN_CFG_PDL off
timed(bwd, "theta + kframe, no programmatic dependent launch")
L.loans_stn_configure(1, 1)
