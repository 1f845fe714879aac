#!/usr/bin/env python
"""bench.py -- localizer crops/sec of the STN crop path (fwd+bwd) on B200, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's numpy CPU path, restated, on the host cores)

A step = one forward + one backward of the fused path over one batch of synthetic frames of BASELINE.json's
configs[1] (cfg2: batch 64, 3x224x224 -> 64x64, fp32, gx produced), per GPU (weak scaling: the path shards by
batch with no collective).  The headline runs the path as LoANs ships it -- rotation_dropout(ratio=0.0) in front of
the grid, i.e. axis-aligned crops (reference sheep/sheep_localizer.py:61); the same steps with a general affine
theta (no dropout node) are timed in the same run and reported under "variants".  `value` is timed on the device with inputs resident in
HBM, the steps replayed from CUDA graphs (2 kernel launches per step); `e2e` is the same work through the public
operators with pinned HOST buffers, copies inside the timed region.  Nothing here reads /root/reference.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "localizer crops/sec (STN fwd+bwd)"
UNIT = "crops/s"
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--no-gx", action="store_true", help="frames do not require grad (what the LoANs step needs)")
    ap.add_argument("--rotation-ratio", default="shipped",
                    help="'shipped' (default): rotation_dropout(ratio=0.0) in front of the grid, as LoANs always calls it "
                         "(sheep/sheep_localizer.py:61) -> axis-aligned crops; 'none': no dropout node, general affine theta; "
                         "or a number used as the test-mode mask value")
    ap.add_argument("--no-variants", action="store_true", help="skip the second (general-affine / as-shipped) timing")
    ap.add_argument("--tma-forward", action="store_true", help="opt into the TMA-staged forward kernel (axis-aligned crops)")
    ap.add_argument("--no-pdl", action="store_true", help="plain launches instead of programmatic dependent launch (A/B)")
    ap.add_argument("--band", default="auto", choices=["auto", "on", "off"], help="band backward kernel: library default / always / never")
    ap.add_argument("--no-cudnn", action="store_true", help="skip the cuDNN comparison arm (the reference's GPU path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU work per worker for the cpu_baseline leg")
    ap.add_argument("--ref-frames", type=int, default=64,
                    help="--impl reference: frames of the batch one step works through (bounded sample for the big configs)")
    return ap.parse_args()


def workload_config(wl, need_gx, extra=None):
    cfg = {"workload": "%s: %s" % (wl.name, wl.description), "batch_per_gpu": wl.batch,
           "crops_per_frame": wl.crops_per_frame, "frame": [wl.channels, wl.height, wl.width],
           "crop": [wl.out_h, wl.out_w], "crop_dtype": wl.out_dtype,
           "rotation_dropout_ratio": wl.rotation_ratio, "gx": bool(need_gx)}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ CPU legs
def _cpu_worker(task):
    """Runs in a spawned process: numpy restatement of the reference's CPU path, fwd+bwd, on its own shard."""
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    name, frames, seed, need_gx, seconds, reps = task[:6]
    from loans_b200 import workloads as W
    from oracle import stn_numpy as on
    wl = W.WORKLOADS[name]
    if len(task) > 6:                                  # rotation-dropout ratio of the arm being mirrored
        wl = wl._replace(rotation_ratio=task[6])
    d = W.make_inputs(wl, seed=seed, batch=frames)
    osz = (wl.out_h, wl.out_w)
    mask = 1.0 if wl.rotation_ratio is None else float(wl.rotation_ratio)
    k = wl.crops_per_frame

    def one():
        on.crop_forward(d["x"], d["theta"], osz, mask, k)
        on.crop_backward(d["x"], d["theta"], osz, d["gy"], None, mask, k)

    one()                                              # warm-up
    if reps is None:
        t0 = time.perf_counter()
        one()
        t1 = time.perf_counter() - t0
        reps = max(1, int(math.ceil(seconds / max(t1, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    el = time.perf_counter() - t0
    return frames * k * reps, el


def cpu_baseline(wl_name, need_gx, seconds, ratio=0.0):
    """All host cores, one process each (the numpy path is single-threaded by construction), every process
    working through its own 8-frame shard of the workload for ~`seconds`; rate = sum of per-process rates."""
    import multiprocessing as mp
    from loans_b200 import workloads as W
    wl = W.WORKLOADS[wl_name]
    cores = os.cpu_count() or 1
    frames = max(1, min(8, wl.batch))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(wl_name, frames, 1234 + i, need_gx, seconds, None, ratio) for i in range(cores)], chunksize=1)
    rate = sum(c / t for c, t in res)
    one = max(c / t for c, t in res)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "value_1core": one,
            "sample": "numpy restatement of chainer 4.1.0's CPU sampler/grid + reference rotation dropout (oracle/stn_numpy.py), "
                      "fwd+bwd incl. gx, %d processes x %d-frame shards of %s for ~%.0f s each"
                      % (cores, frames, wl.name, seconds)}


def _ref_worker(conn, name, frames, seed, ratio):
    """Persistent process of the reference arm: builds its shard of the batch ONCE, then runs one fwd+bwd of the numpy
    path per "step" message -- input generation, imports and warm-up stay outside the timed steps."""
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    from loans_b200 import workloads as W
    from oracle import stn_numpy as on
    wl = W.WORKLOADS[name]._replace(rotation_ratio=ratio)
    d = W.make_inputs(wl, seed=seed, batch=frames)
    osz = (wl.out_h, wl.out_w)
    mask = 1.0 if wl.rotation_ratio is None else float(wl.rotation_ratio)
    k = wl.crops_per_frame
    conn.send("ready")
    while conn.recv() == "step":
        on.crop_forward(d["x"], d["theta"], osz, mask, k)
        on.crop_backward(d["x"], d["theta"], osz, d["gy"], None, mask, k)
        conn.send("done")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (chainer is not installable here, so
    its numpy restatement, the oracle port) on all host cores, same config/metric as our arm.  A step is one batch of
    the workload (a bounded slice of it for the big configs), its frames split over one persistent process per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from loans_b200 import workloads as W
    wl = W.WORKLOADS[args.workload]
    rr = args.rotation_ratio                       # the same config as our arm: LoANs' ratio 0.0 unless asked otherwise
    wl = wl._replace(rotation_ratio=0.0 if rr == "shipped" else (None if rr == "none" else float(rr)))
    cores = os.cpu_count() or 1
    frames_step = min(wl.batch, max(cores, args.ref_frames))    # bounded sample: the CPU path is ~2.5 ms per crop per core
    procs = max(1, min(cores, frames_step))
    base, extra = divmod(frames_step, procs)
    shards = [base + (1 if i < extra else 0) for i in range(procs)]
    ctx = mp.get_context("spawn")
    workers = []
    for i, f in enumerate(shards):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(b, wl.name, f, 1234 + i, wl.rotation_ratio), daemon=True)
        pr.start()
        workers.append((pr, a))
    for _, a in workers:
        assert a.recv() == "ready"

    def step():
        for _, a in workers:
            a.send("step")
        for _, a in workers:
            assert a.recv() == "done"

    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    for pr, a in workers:
        a.send("stop")
    for pr, _ in workers:
        pr.join(timeout=10)
    crops = frames_step * wl.crops_per_frame * args.steps
    value = crops / el
    sample = ("each step = %d of the %d frames of one %s batch, split over %d persistent processes (inputs built once, "
              "outside the timed steps); numpy restatement of the reference's CPU path (oracle/stn_numpy.py), fwd+bwd "
              "incl. gx" % (frames_step, wl.batch, wl.name, procs))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, True),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []           # (t, sm_mhz, reasons_mask, power_w)
        self.windows = []
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:          # NVML missing: report that instead of clocks
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = float("nan")
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(2.0)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % getattr(self, "err", "?")]}
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        pool = inside if inside else self.samples
        sms = sorted(s[1] for s in pool)
        mask = 0
        for s in pool:
            mask |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if mask & bit]
        pw = [s[3] for s in pool if s[3] == s[3]]
        return {"sm_mhz": sms[len(sms) // 2] if sms else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(pool), "samples_in_timed_regions": len(inside), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from loans_b200 import _lib
    from loans_b200 import workloads as W
    from loans_b200.functions import stn_crop

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = W.WORKLOADS[args.workload]
    rr = args.rotation_ratio
    wl = wl._replace(rotation_ratio=0.0 if rr == "shipped" else (None if rr == "none" else float(rr)))
    need_gx = not args.no_gx
    steps, warm = args.steps, max(args.warmup, 3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, need_gx, args.cpu_seconds, wl.rotation_ratio)      # before CUDA is touched: plain host work

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if args.tma_forward:
        _lib.tma_forward(True)
    if args.no_pdl:
        _lib.pdl(False)
    if os.environ.get("STN_THETA_ONLY_KERNEL") == "0":      # A/B: gx == NULL through the two-role kernel
        _lib.check(L.loans_stn_configure(11, 0), "loans_stn_configure")
    if args.band != "auto":
        _lib.band_backward(args.band == "on")
    sampler = ClockSampler(local_rank)
    sampler.start()

    B, K, C, H, Wd, oH, oW = wl.batch, wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
    N = B * K
    bf16 = wl.out_dtype == "bf16"
    ydt = torch.bfloat16 if bf16 else torch.float32
    dt_code = _lib.BF16 if bf16 else _lib.F32
    mask01 = 1.0 if wl.rotation_ratio is None else float(wl.rotation_ratio)   # train-mode draw at ratio 0.0 is always 0
    fwd_bytes, bwd_bytes = W.algorithmic_bytes(wl, need_gx=need_gx)

    # ---- input sets: rotated so that consecutive steps never find their inputs in the 126 MB L2
    set_bytes = 4 * B * C * H * Wd * (2 if need_gx else 1) + N * C * oH * oW * (2 if bf16 else 4) * 2 + N * 2 * oH * oW * 4
    S = int(min(16, max(4, math.ceil(3.0 * L2_BYTES / set_bytes))))
    sets = []
    for s in range(S):
        d = W.make_inputs(wl, seed=1234 + 1000 * rank + s)
        e = {"x": torch.from_numpy(d["x"]).to(dev), "theta": torch.from_numpy(d["theta"]).to(dev),
             "gy": torch.from_numpy(d["gy"]).to(dev).to(ydt),
             "y": torch.empty((N, C, oH, oW), dtype=ydt, device=dev),
             "grid": torch.empty((N, 2, oH, oW), dtype=torch.float32, device=dev),
             "gtheta": torch.empty((N, 2, 3), dtype=torch.float32, device=dev),
             "gx": torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev) if need_gx else None}
        if s == 0:
            host0 = d
        sets.append(e)

    def p(t):
        return None if t is None else t.data_ptr()

    def fwd(e):
        _lib.check(L.loans_stn_crop_fwd(p(e["x"]), p(e["theta"]), float(mask01), p(e["y"]), p(e["grid"]), N, K, C, H, Wd, oH, oW,
                                        dt_code, torch.cuda.current_stream().cuda_stream), "crop_fwd")

    def bwd(e):
        _lib.check(L.loans_stn_crop_bwd(p(e["x"]), p(e["theta"]), float(mask01), p(e["gy"]), None, p(e["gtheta"]), p(e["gx"]), None,
                                        N, K, C, H, Wd, oH, oW, dt_code, torch.cuda.current_stream().cuda_stream), "crop_bwd")

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    n0 = _lib.launch_count()
    fwd(sets[0]); bwd(sets[0])
    launches_per_step = _lib.launch_count() - n0
    torch.cuda.synchronize()
    g_all = capture(lambda: [(fwd(e), bwd(e)) for e in sets])
    g_one = [capture(lambda e=e: (fwd(e), bwd(e))) for e in sets]
    g_fwd = capture(lambda: [fwd(e) for e in sets])
    g_bwd = capture(lambda: [bwd(e) for e in sets])

    def run_steps(k):
        q, r = divmod(k, S)
        for _ in range(q):
            g_all.replay()
        for s in range(r):
            g_one[s].replay()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sampler.window(t0, t1)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    # ---- warm-up: W steps, and at least ~0.3 s of work so that the clocks have ramped
    run_steps(warm)
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.3 and not os.environ.get("STN_BENCH_NO_RAMP"):
        run_steps(S * 8)
        torch.cuda.synchronize()

    # ---- the timed region: exactly K steps
    ms_total = timed(lambda: run_steps(steps))
    ms_step = ms_total / steps
    value = world * N * steps / (ms_total * 1e-3)

    # ---- per-kernel durations (the same launches, forward-only and backward-only graphs)
    reps = max(1, steps // S)
    ms_f = timed(lambda: [g_fwd.replay() for _ in range(reps)]) / (reps * S)
    ms_b = timed(lambda: [g_bwd.replay() for _ in range(reps)]) / (reps * S)

    # ---- the other theta regime, same buffers: general affine (mask 1) if the headline is as-shipped, and vice versa
    variants = {}
    if not args.no_variants:
        main_mask = mask01
        alt_mask = 1.0 if main_mask == 0.0 else 0.0
        mask01 = alt_mask
        a_all = capture(lambda: [(fwd(e), bwd(e)) for e in sets])
        a_fwd = capture(lambda: [fwd(e) for e in sets])
        a_bwd = capture(lambda: [bwd(e) for e in sets])
        mask01 = main_mask
        for _ in range(3):
            a_all.replay()
        reps_a = max(1, steps // S)
        ms_a = timed(lambda: [a_all.replay() for _ in range(reps_a)]) / (reps_a * S)
        ms_af = timed(lambda: [a_fwd.replay() for _ in range(reps_a)]) / (reps_a * S)
        ms_ab = timed(lambda: [a_bwd.replay() for _ in range(reps_a)]) / (reps_a * S)
        variants["general_affine" if alt_mask == 1.0 else "as_shipped_axis_aligned"] = {
            "mask01": alt_mask, "value": world * N / (ms_a * 1e-3), "unit": UNIT, "us_per_step": ms_a * 1e3,
            "fwd_us": ms_af * 1e3, "bwd_us": ms_ab * 1e3,
            "whole_step_frac": (fwd_bytes + bwd_bytes) / (ms_a * 1e-3) / 1e9}

    # ---- SURVEY.md section 8(f) rows built so far, timed on the same buffers: prepare_images (rank 1), corner points (rank 2)
    next_rows = {}
    if not args.no_variants:
        prep_out = torch.empty_like(sets[0]["x"])

        def prep(e):                                   # output into the set's own gx buffer: rotating, so the writes reach DRAM
            _lib.check(L.loans_stn_prepare_images(p(e["x"]), 255.0, p(e["gx"] if need_gx else prep_out), B, C, H, Wd,
                                                  torch.cuda.current_stream().cuda_stream), "prepare_images")
        if C == 3:
            prep(sets[0])
            g_prep = capture(lambda: [prep(e) for e in sets])
            for _ in range(3):
                g_prep.replay()
            reps_p = max(1, steps // S)
            ms_p = timed(lambda: [g_prep.replay() for _ in range(reps_p)]) / (reps_p * S)
            pb = 8 * B * C * H * Wd
            next_rows["prepare_images"] = {"us": ms_p * 1e3, "frames_per_s": world * B / (ms_p * 1e-3),
                                           "algorithmic_bytes": pb, "achieved_gbs": pb / (ms_p * 1e-3) / 1e9,
                                           "what": "SheepLocalizer.prepare_images as one kernel (uint8 quantise, RGB->BGR, mean), x*255 folded in"}
        cor = torch.empty((N, 2, 2, 2), dtype=torch.float32, device=dev)
        gcor = torch.randn((N, 2, 2, 2), dtype=torch.float32, device=dev)

        def step_corners(e):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.loans_stn_crop_fwd_corners(p(e["x"]), p(e["theta"]), float(mask01), p(e["y"]), p(cor), N, K, C, H, Wd, oH, oW,
                                                    dt_code, st), "crop_fwd_corners")
            _lib.check(L.loans_stn_crop_bwd_corners(p(e["x"]), p(e["theta"]), float(mask01), p(e["gy"]), p(gcor), p(e["gtheta"]),
                                                    p(e["gx"]), N, K, C, H, Wd, oH, oW, dt_code, st), "crop_bwd_corners")
        step_corners(sets[0])
        g_cor = capture(lambda: [step_corners(e) for e in sets])
        for _ in range(3):
            g_cor.replay()
        reps_c = max(1, steps // S)
        ms_c = timed(lambda: [g_cor.replay() for _ in range(reps_c)]) / (reps_c * S)
        if C == 3:
            yg = torch.empty((N, 1, oH, oW), dtype=ydt, device=dev)
            gyg = torch.randn((N, 1, oH, oW), dtype=torch.float32, device=dev).to(ydt)

            def step_gray(e):
                st = torch.cuda.current_stream().cuda_stream
                _lib.check(L.loans_stn_crop_fwd_ex(p(e["x"]), p(e["theta"]), float(mask01), p(yg), None, p(cor), _lib.FLAG_GRAY,
                                                   N, K, C, H, Wd, oH, oW, dt_code, st), "crop_fwd_ex")
                _lib.check(L.loans_stn_crop_bwd_ex(p(e["x"]), p(e["theta"]), float(mask01), p(gyg), None, p(gcor), p(e["gtheta"]),
                                                   p(e["gx"]), None, _lib.FLAG_GRAY, N, K, C, H, Wd, oH, oW, dt_code, st), "crop_bwd_ex")
            step_gray(sets[0])
            g_gray = capture(lambda: [step_gray(e) for e in sets])
            for _ in range(3):
                g_gray.replay()
            ms_g = timed(lambda: [g_gray.replay() for _ in range(reps_c)]) / (reps_c * S)
            next_rows["grayscale_corners"] = {"us_per_step": ms_g * 1e3, "value": world * N / (ms_g * 1e-3), "unit": UNIT,
                                              "what": "fwd+bwd with the localizer's grayscale epilogue fused (1-channel crops "
                                                      "and gy) and corner points"}
        next_rows["corner_points"] = {"us_per_step": ms_c * 1e3, "value": world * N / (ms_c * 1e-3), "unit": UNIT,
                                      "what": "fwd+bwd with points reduced to the grid's four corners (no dense grid written), "
                                              "corner gradient folded into gtheta"}

    # ---- the reference's GPU path on the same inputs: cuDNN's spatial-transformer kernels (what chainer calls on a GPU)
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cudnn and K == 1 and not bf16 and need_gx:
        try:
            from baseline.cudnn_stn import time_cudnn
            us_c, ver, (y_c, gx_c, gt_c) = time_cudnn(wl, sets, float(mask01), max(1, steps // S), dev)
            fwd(sets[0]); bwd(sets[0])
            torch.cuda.synchronize()

            def rel(a, b):
                return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
            gpu_ref = {"impl": "cuDNN %d cudnnSpatialTfGridGenerator/Sampler Forward+Backward + chainer's grid layout copy "
                               "(the kernels chainer 4.1.0 runs for this path on a GPU), same inputs, CUDA-graph replay" % ver,
                       "us_per_step": us_c, "value": N / (us_c * 1e-6), "unit": UNIT,
                       "ours_vs_cudnn_speedup": us_c / (ms_step * 1e3),
                       "max_rel_diff_vs_ours": {"y": rel(sets[0]["y"].float(), y_c), "gx": rel(sets[0]["gx"], gx_c),
                                                "gtheta": rel(sets[0]["gtheta"], gt_c)}}
        except Exception as e:            # comparison arm only: report why it is missing
            gpu_ref = {"unavailable": repr(e)[:300]}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s%s" % (wl.name, "" if need_gx else "_nogx"), {}).get("bwd_dram_bytes")
    # ---- the DRAM write rate of this GPU, measured here: zero-fill of the rotating gx buffers (plain torch fill kernels, a
    # calibration like MEASURED_PEAKS.json's copy, not part of the path).  The backward is write-dominated (gx is dense): this is
    # what the same bytes cost when nothing but the stores is done, in the same graph harness, launch included.
    write_cal = None
    if need_gx:
        g_fill = capture(lambda: [e["gx"].zero_() for e in sets])
        for _ in range(3):
            g_fill.replay()
        reps_w = max(1, steps // S)
        ms_w = timed(lambda: [g_fill.replay() for _ in range(reps_w)]) / (reps_w * S)
        gx_bytes = 4 * B * C * H * Wd
        write_gbs = gx_bytes / (ms_w * 1e-3) / 1e9
        read_bytes = bwd_bytes - gx_bytes
        floor_us = (gx_bytes / write_gbs + read_bytes / peak) / 1e3
        write_cal = {"write_gbs_measured": write_gbs, "fill_us": ms_w * 1e3, "bwd_written_bytes": gx_bytes,
                     "bwd_read_bytes": read_bytes, "bwd_dram_floor_us": floor_us,
                     "bwd_frac_of_dram_floor": floor_us / (ms_b * 1e3),
                     "note": "floor = gx bytes at the measured pure-write rate + the algorithmic read bytes at the copy peak"}
    ach_b = bwd_bytes / (ms_b * 1e-3) / 1e9
    ach_f = fwd_bytes / (ms_f * 1e-3) / 1e9
    ach_s = (fwd_bytes + bwd_bytes) / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "backward launch: stn_bwd_band_kernel (row bands or CTA bands) where the band backward is taken (mask01 == 0, one crop per frame; rule in launch_crop_bwd_band), else stn_bwd_kernel (gx role + cluster-reduced theta role)",
                "achieved": ach_b, "peak": peak, "unit": "GB/s", "frac": ach_b / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes, "avg_launch_us": ms_b * 1e3,
                "write_bound": write_cal,
                "fwd_kernel": {"achieved": ach_f, "frac": ach_f / peak, "algorithmic_bytes_per_launch": fwd_bytes,
                               "avg_launch_us": ms_f * 1e3},
                "whole_step": {"achieved": ach_s, "frac": ach_s / peak, "algorithmic_bytes": fwd_bytes + bwd_bytes,
                               "us": ms_step * 1e3}}

    # ---- e2e: public operators, pinned host buffers, H2D + D2H inside the timed region, every step
    e2e = None
    if not args.no_e2e:
        hx = torch.from_numpy(host0["x"]).pin_memory()
        hth = torch.from_numpy(host0["theta"]).pin_memory()
        hgy = torch.from_numpy(host0["gy"]).to(ydt).pin_memory()
        ry = torch.empty((N, C, oH, oW), dtype=ydt).pin_memory()
        rgrid = torch.empty((N, 2, oH, oW), dtype=torch.float32).pin_memory()
        rgt = torch.empty((N, 2, 3), dtype=torch.float32).pin_memory()
        rgx = torch.empty((B, C, H, Wd), dtype=torch.float32).pin_memory() if need_gx else None
        dx = torch.empty_like(sets[0]["x"]).requires_grad_(need_gx)
        dth = torch.empty_like(sets[0]["theta"]).requires_grad_(True)
        dgy = torch.empty_like(sets[0]["gy"])
        h2d = hx.numel() * 4 + hth.numel() * 4 + hgy.numel() * hgy.element_size()
        d2h = ry.numel() * ry.element_size() + rgrid.numel() * 4 + rgt.numel() * 4 + (rgx.numel() * 4 if need_gx else 0)

        def e2e_step():
            with torch.no_grad():
                dx.copy_(hx, non_blocking=True)
                dth.copy_(hth, non_blocking=True)
                dgy.copy_(hgy, non_blocking=True)
            dx.grad = None
            dth.grad = None
            rois, points = stn_crop(dx, dth, (oH, oW), mask01=mask01, crops_per_frame=K, out_dtype=ydt)
            torch.autograd.backward([rois], [dgy])
            ry.copy_(rois.detach(), non_blocking=True)
            rgrid.copy_(points.detach(), non_blocking=True)
            rgt.copy_(dth.grad, non_blocking=True)
            if need_gx:
                rgx.copy_(dx.grad, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the step's results are on the host

        e2e_steps = max(3, min(steps, int(2.0 / max(1e-4, (h2d + d2h) / 20e9))))
        for _ in range(3):
            e2e_step()
        ms_serial = timed(lambda: [e2e_step() for _ in range(e2e_steps)])
        # the same steps through the host-buffer pipeline (H2D / compute / D2H on three streams, two buffer sets):
        # every step still uploads its inputs from pinned memory and downloads all its results, inside the timed region
        from loans_b200.pipeline import HostCropPipeline
        pipe = HostCropPipeline(B, C, H, Wd, (oH, oW), crops_per_frame=K, need_gx=need_gx, out_dtype=ydt, depth=2, device=dev)
        outs = [{"y": torch.empty_like(ry).pin_memory(), "grid": torch.empty_like(rgrid).pin_memory(),
                 "gtheta": torch.empty_like(rgt).pin_memory(), "gx": torch.empty_like(rgx).pin_memory() if need_gx else None}
                for _ in range(2)]
        for i in range(4):
            pipe.submit(hx, hth, hgy, outs[i % 2], mask01=mask01)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.s_in)
        for i in range(e2e_steps):
            pipe.submit(hx, hth, hgy, outs[i % 2], mask01=mask01)
        e1.record(pipe.s_out)
        pipe.drain()
        torch.cuda.synchronize()
        sampler.window(t0, time.perf_counter())
        ms_e = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
            dist.barrier()
        ok = bool(torch.equal(outs[(e2e_steps - 1) % 2]["y"], ry))            # pipeline and plain call agree bit for bit
        e2e = {"value": world * N * e2e_steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps, "ms_per_step": ms_e / e2e_steps,
               "api": "loans_b200.pipeline.HostCropPipeline (pinned host tensors in and out; C-ABI fwd+bwd on a compute stream, "
                      "H2D and D2H on their own streams, two buffer sets)",
               "serial": {"value": world * N * e2e_steps / (ms_serial * 1e-3), "ms_per_step": ms_serial / e2e_steps,
                          "api": "loans_b200.functions.stn_crop + autograd backward, copy-in / run / copy-out one step at a time"},
               "matches_serial_result": ok}

    sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": workload_config(wl, need_gx, {
                    "l2": "rotating %d distinct input/output sets (%.0f MB each, %.0f MB total > 126 MB L2)"
                          % (S, set_bytes / 1e6, S * set_bytes / 1e6),
                    "launch": "CUDA-graph replay of the C-ABI calls loans_stn_crop_fwd + loans_stn_crop_bwd, "
                              + ("plain launches" if args.no_pdl else "programmatic dependent launch (griddepcontrol) between consecutive kernels"),
                    "band_backward": args.band}),
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches_per_step * steps,
                "roofline": roofline, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "variants": variants, "next_rows": next_rows}
        for v in variants.values():
            v["whole_step_frac"] = v["whole_step_frac"] / peak
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
