#!/usr/bin/env python
"""bench.py -- localizer crops/sec of the STN crop path (fwd+bwd) on B200, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's numpy CPU path, restated, on the host cores)

A step = one forward + one backward of the fused path over one batch of synthetic frames of BASELINE.json's
configs[1] (cfg2: batch 64, 3x224x224 -> 64x64, fp32, gx produced), per GPU (weak scaling: the path shards by
batch with no collective).  The headline runs the path as LoANs ships it -- rotation_dropout(ratio=0.0) in front of
the grid, i.e. axis-aligned crops (reference sheep/sheep_localizer.py:61).  In the same run, on the same GPU:

  variants.general_affine   the same steps with a general affine theta (no dropout node): BASELINE cfg2 as literally worded;
  variants.drop_in          the same steps written as the reference's THREE public calls (rotation_dropout ->
                            spatial_transformer_grid -> spatial_transformer_sampler + autograd backward), CUDA-graphed;
  configs                   every other BASELINE config (cfg1, cfg3, cfg4, cfg5, cfg2 without gx) with its own roofline fractions;
  cfg5_sharded              BASELINE configs[4]: global batch 1024 sharded over the N ranks, fwd+bwd per shard plus the step's
                            one collective (NCCL mean all-reduce of a localizer-sized fp32 gradient bucket) in the timed step;
  floor                     what kernels of this launch shape cost before any STN arithmetic (empty graph nodes, two dependent
                            DRAM round trips, the same bytes streamed by a plain copy-like kernel).

`value` is timed on the device with inputs resident in HBM, the steps replayed from CUDA graphs (2 kernel launches per
step); `e2e` is the same work through the host-buffer API with pinned HOST buffers, copies inside the timed region.
Nothing here reads /root/reference.
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
LOCALIZER_PARAMS = 12_592_902        # ResNet-18 trunk 12,589,824 + Linear(512,6) 3,078 (reference sheep/resnet.py, sheep/sheep_localizer.py:23-28)
# CPU legs (cpu_baseline of our arm, --impl reference): a TASK is one fwd+bwd of the numpy path over a worker's own shard
# of CPU_TASK_FRAMES frames; identical in both legs.  The reference arm's step is CPU_TASKS_PER_CORE x cores tasks handed
# out one at a time to whichever worker process is free (>= ~100 ms per step, so the step barrier costs a few per cent)
CPU_TASK_FRAMES = {"cfg1": 2, "cfg2": 2, "cfg3": 1, "cfg4": 1, "cfg5": 2}
CPU_TASKS_PER_CORE = {"cfg1": 24, "cfg2": 24, "cfg3": 12, "cfg4": 1, "cfg5": 24}


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
    ap.add_argument("--no-variants", action="store_true", help="skip the general-affine / drop-in / next-rows timings")
    ap.add_argument("--no-configs", action="store_true", help="skip the block with the other BASELINE configs")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the batch-sharded cfg5 block (with the gradient all-reduce)")
    ap.add_argument("--no-floor", action="store_true", help="skip the launch / latency / streaming floor probes")
    ap.add_argument("--no-pdl", action="store_true", help="plain launches instead of programmatic dependent launch (A/B)")
    ap.add_argument("--band", default="auto", choices=["auto", "on", "off"], help="band backward kernel: library default / always / never")
    ap.add_argument("--no-cudnn", action="store_true", help="skip the cuDNN comparison arm (the reference's GPU path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU work per worker for the cpu_baseline leg")
    ap.add_argument("--ref-tasks", type=int, default=0,
                    help="--impl reference: tasks per core and step (0: CPU_TASKS_PER_CORE, >= ~100 ms per step)")
    return ap.parse_args()


def workload_config(wl, need_gx):
    """Names the workload; the SAME keys in both arms (measurement-harness details go to the line's `harness` key)."""
    return {"workload": "%s: %s" % (wl.name, wl.description), "batch_per_gpu": wl.batch,
            "crops_per_frame": wl.crops_per_frame, "frame": [wl.channels, wl.height, wl.width],
            "crop": [wl.out_h, wl.out_w], "crop_dtype": wl.out_dtype,
            "rotation_dropout_ratio": wl.rotation_ratio, "gx": bool(need_gx)}


def pick_ratio(args, wl):
    rr = args.rotation_ratio                        # LoANs' ratio 0.0 unless asked otherwise
    return wl._replace(rotation_ratio=0.0 if rr == "shipped" else (None if rr == "none" else float(rr)))


# ------------------------------------------------------------------------------------------------ CPU legs
def _cpu_setup(name, frames, seed, ratio):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    from loans_b200 import workloads as W
    from oracle import stn_numpy as on
    wl = W.WORKLOADS[name]._replace(rotation_ratio=ratio)
    d = W.make_inputs(wl, seed=seed, batch=frames)
    osz = (wl.out_h, wl.out_w)
    mask = 1.0 if wl.rotation_ratio is None else float(wl.rotation_ratio)
    k = wl.crops_per_frame

    def one():
        on.crop_forward(d["x"], d["theta"], osz, mask, k)
        on.crop_backward(d["x"], d["theta"], osz, d["gy"], None, mask, k)
    return one, frames * k


def _cpu_worker(task):
    """Runs in a spawned process: numpy restatement of the reference's CPU path, fwd+bwd, on its own shard, free-running."""
    name, frames, seed, seconds, ratio = task
    one, crops = _cpu_setup(name, frames, seed, ratio)
    one()                                              # warm-up
    t0 = time.perf_counter()
    one()
    t1 = time.perf_counter() - t0
    reps = max(1, int(math.ceil(seconds / max(t1, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    return crops * reps, time.perf_counter() - t0


def cpu_baseline(wl_name, seconds, ratio, frames=None):
    """All host cores, one process each (the numpy path is single-threaded by construction), every process working
    through its own shard of the workload for ~`seconds`; rate = sum of the per-process rates."""
    import multiprocessing as mp
    from loans_b200 import workloads as W
    wl = W.WORKLOADS[wl_name]
    cores = os.cpu_count() or 1
    frames = frames or CPU_TASK_FRAMES[wl_name]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(wl_name, frames, 1234 + i, seconds, ratio) for i in range(cores)], chunksize=1)
    rate = sum(c / t for c, t in res)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "value_1core": max(c / t for c, t in res),
            "sample": "numpy restatement of chainer 4.1.0's CPU sampler/grid + reference rotation dropout (oracle/stn_numpy.py), "
                      "fwd+bwd incl. gx, %d processes x %d-frame shards of %s, free-running for ~%.0f s each"
                      % (cores, frames, wl.name, seconds)}


def _ref_worker(conn, counter, name, frames, seed, ratio):
    """Persistent process of the reference arm: builds its shard ONCE, then, per "step" message, runs tasks (one fwd+bwd of
    the numpy path over the shard) for as long as the step's shared task counter hands it one -- input generation, imports
    and warm-up stay outside the timed steps."""
    one, _ = _cpu_setup(name, frames, seed, ratio)
    one()
    conn.send("ready")
    while conn.recv() == "step":
        done = 0
        while True:
            with counter.get_lock():
                left = counter.value
                if left > 0:
                    counter.value = left - 1
            if left <= 0:
                break
            one()
            done += 1
        conn.send(done)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (chainer is not installable here, so its
    numpy restatement, the oracle port) on all host cores, same config / metric as our arm.  A step = CPU_TASKS_PER_CORE x
    cores tasks (a task = fwd+bwd over one CPU_TASK_FRAMES-frame shard, the same task the cpu_baseline leg of our arm
    repeats), handed out one at a time to whichever of the persistent worker processes (one per core) is free."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from loans_b200 import workloads as W
    wl = pick_ratio(args, W.WORKLOADS[args.workload])
    cores = os.cpu_count() or 1
    frames = CPU_TASK_FRAMES[wl.name]
    tasks = cores * (args.ref_tasks if args.ref_tasks > 0 else CPU_TASKS_PER_CORE[wl.name])
    ctx = mp.get_context("spawn")
    counter = ctx.Value("i", 0)
    workers = []
    for i in range(cores):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(b, counter, wl.name, frames, 1234 + i, wl.rotation_ratio), daemon=True)
        pr.start()
        workers.append((pr, a))
    for _, a in workers:
        assert a.recv() == "ready"

    def step():
        with counter.get_lock():
            counter.value = tasks
        for _, a in workers:
            a.send("step")
        assert sum(a.recv() for _, a in workers) == tasks

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
    crops_step = tasks * frames * wl.crops_per_frame
    value = crops_step * args.steps / el
    sample = ("each step = %d tasks (%d crops) shared out to %d persistent processes; a task = one fwd+bwd incl. gx of the numpy "
              "restatement of the reference's CPU path (oracle/stn_numpy.py) over a %d-frame shard of %s; shards built once, "
              "outside the timed steps" % (tasks, crops_step, cores, frames, wl.name))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, True),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
class Harness(object):
    """Device, distributed world, timing.  timed(fn): barrier + device sync on both sides, CUDA events on the launch stream,
    max over ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local_rank)
        self.sampler.start()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, v):
        if self.world == 1:
            return [float(v)]
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def timed(self, fn, reduce=True):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.sampler.window(t0, t1)
        ms = e0.elapsed_time(e1)
        if reduce and self.world > 1:
            ms = self.max_over_ranks(ms)
            self.dist.barrier()
        return ms

    def capture(self, fn):
        g = self.torch.cuda.CUDAGraph()
        with self.torch.cuda.graph(g):
            fn()
        return g


def ptr(t):
    return None if t is None else t.data_ptr()


class PathBench(object):
    """One workload resident in HBM as S rotating input/output sets (so that consecutive steps never find their inputs in
    the 126 MB L2), its C-ABI forward / backward calls and their CUDA graphs."""

    def __init__(self, hz, wl, need_gx, seed=1234, host_set0=False, batch=None, min_sets=4):
        import numpy as np
        from loans_b200 import _lib
        from loans_b200 import workloads as W
        torch = hz.torch
        self.hz, self.wl, self.need_gx, self.lib, self._lib = hz, wl, need_gx, _lib.lib(), _lib
        dev = hz.dev
        self.B = wl.batch if batch is None else batch
        self.K, self.C, self.H, self.Wd, self.oH, self.oW = wl.crops_per_frame, wl.channels, wl.height, wl.width, wl.out_h, wl.out_w
        B, K, C, H, Wd, oH, oW = self.B, self.K, self.C, self.H, self.Wd, self.oH, self.oW
        self.N = N = B * K
        self.bf16 = wl.out_dtype == "bf16"
        self.ydt = torch.bfloat16 if self.bf16 else torch.float32
        self.dt_code = _lib.BF16 if self.bf16 else _lib.F32
        self.mask01 = 1.0 if wl.rotation_ratio is None else float(wl.rotation_ratio)   # train-mode draw at ratio 0.0 is always 0
        self.fwd_bytes, self.bwd_bytes = W.algorithmic_bytes(wl, need_gx=need_gx, batch=B)
        self.set_bytes = 4 * B * C * H * Wd * (2 if need_gx else 1) + N * C * oH * oW * (2 if self.bf16 else 4) * 2 + N * 2 * oH * oW * 4
        self.S = S = int(min(16, max(min_sets, math.ceil(3.0 * L2_BYTES / self.set_bytes))))
        self.sets, self.host0 = [], None
        gen = torch.Generator(device=dev)
        rotate = wl.rotation_ratio is None
        for s in range(S):
            sd = seed + 1000 * hz.rank + s
            if host_set0 and s == 0:
                d = W.make_inputs(wl, seed=sd, batch=B)
                self.host0 = d
                x, gy = torch.from_numpy(d["x"]).to(dev), torch.from_numpy(d["gy"]).to(dev)
                theta = torch.from_numpy(d["theta"]).to(dev)
            else:
                # same distributions as workloads.make_inputs (uniform [0,1) frames, standard-normal gy), drawn on the device;
                # theta from the workload's own generator (per-frame base boxes + K jitters for cfg4)
                gen.manual_seed(sd)
                x = torch.rand((B, C, H, Wd), dtype=torch.float32, device=dev, generator=gen)
                gy = torch.randn((N, C, oH, oW), dtype=torch.float32, device=dev, generator=gen)
                theta = torch.from_numpy(_theta_only(W, wl, sd, B, rotate)).to(dev)
            e = {"x": x, "theta": theta, "gy": gy.to(self.ydt),
                 "y": torch.empty((N, C, oH, oW), dtype=self.ydt, device=dev),
                 "grid": torch.empty((N, 2, oH, oW), dtype=torch.float32, device=dev),
                 "gtheta": torch.empty((N, 2, 3), dtype=torch.float32, device=dev),
                 "gx": torch.empty((B, C, H, Wd), dtype=torch.float32, device=dev) if need_gx else None}
            self.sets.append(e)
        self.np = np
        self.kernels = None

    def stream(self):
        return self.hz.torch.cuda.current_stream().cuda_stream

    def fwd(self, e):
        self._lib.check(self.lib.loans_stn_crop_fwd(ptr(e["x"]), ptr(e["theta"]), float(self.mask01), ptr(e["y"]), ptr(e["grid"]),
                                                    self.N, self.K, self.C, self.H, self.Wd, self.oH, self.oW, self.dt_code,
                                                    self.stream()), "crop_fwd")

    def bwd(self, e):
        self._lib.check(self.lib.loans_stn_crop_bwd(ptr(e["x"]), ptr(e["theta"]), float(self.mask01), ptr(e["gy"]), None,
                                                    ptr(e["gtheta"]), ptr(e["gx"]), None, self.N, self.K, self.C, self.H, self.Wd,
                                                    self.oH, self.oW, self.dt_code, self.stream()), "crop_bwd")

    def build_graphs(self, per_set=True):
        hz, sets = self.hz, self.sets
        n0 = self._lib.launch_count()
        self.fwd(sets[0])
        kf = self._lib.last_kernel()
        self.bwd(sets[0])
        kb = self._lib.last_kernel()
        self.launches_per_step = self._lib.launch_count() - n0
        self.kernels = {"fwd": kf, "bwd": kb}
        hz.torch.cuda.synchronize()
        self.g_all = hz.capture(lambda: [(self.fwd(e), self.bwd(e)) for e in sets])
        self.g_one = [hz.capture(lambda e=e: (self.fwd(e), self.bwd(e))) for e in sets] if per_set else None
        self.g_fwd = hz.capture(lambda: [self.fwd(e) for e in sets])
        self.g_bwd = hz.capture(lambda: [self.bwd(e) for e in sets])
        return self

    def run_steps(self, k):
        q, r = divmod(k, self.S)
        for _ in range(q):
            self.g_all.replay()
        for s in range(r):
            self.g_one[s].replay()

    def per_launch(self, graph, steps, at_least=False):
        """Average duration of one of the S launches (or steps) a graph holds, over ~`steps` of them (`at_least`: rounded up)."""
        reps = max(1, -(-steps // self.S) if at_least else steps // self.S)
        return self.hz.timed(lambda: [graph.replay() for _ in range(reps)]) / (reps * self.S)

    def fractions(self, ms_step, ms_f, ms_b, peak):
        fb, bb = self.fwd_bytes, self.bwd_bytes
        return {"us_per_step": ms_step * 1e3, "fwd_us": ms_f * 1e3, "bwd_us": ms_b * 1e3,
                "value": self.hz.world * self.N / (ms_step * 1e-3), "unit": UNIT,
                "whole_step_frac": (fb + bb) / (ms_step * 1e-3) / 1e9 / peak,
                "fwd_frac": fb / (ms_f * 1e-3) / 1e9 / peak, "bwd_frac": bb / (ms_b * 1e-3) / 1e9 / peak,
                "algorithmic_bytes": {"fwd": fb, "bwd": bb}, "kernels": self.kernels, "sets": self.S}

    def free(self):
        self.sets = None
        self.g_all = self.g_one = self.g_fwd = self.g_bwd = None
        self.hz.torch.cuda.synchronize()
        self.hz.torch.cuda.empty_cache()


def _theta_only(W, wl, seed, b, rotate):
    """theta of workloads.make_inputs without generating (and throwing away) the frames: same generator calls for theta."""
    import numpy as np
    rng = np.random.default_rng(seed + 7_000_000)
    k = wl.crops_per_frame
    n = b * k
    if k == 1:
        return W.make_theta(rng, n, rotate=rotate)
    theta = np.repeat(W.make_theta(rng, b, rotate=rotate), k, axis=0)
    js = rng.uniform(0.8, 1.25, (n, 2)).astype(np.float32)
    jt = rng.uniform(-0.15, 0.15, (n, 2)).astype(np.float32)
    theta[:, 0, 0] *= js[:, 0]
    theta[:, 1, 1] *= js[:, 1]
    theta[:, 0, 2] += jt[:, 0]
    theta[:, 1, 2] += jt[:, 1]
    return theta.astype(np.float32)


def warm_up(hz, pb, warm):
    pb.run_steps(warm)
    hz.torch.cuda.synchronize()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.3 and not os.environ.get("STN_BENCH_NO_RAMP"):      # at least ~0.3 s of work: clocks ramped
        pb.run_steps(pb.S * 8)
        hz.torch.cuda.synchronize()


def drop_in_variant(hz, pb, steps, peak):
    """The headline's steps written as the reference's three public calls + autograd backward, captured into CUDA graphs."""
    torch = hz.torch
    from loans_b200.functions import rotation_dropout, spatial_transformer_grid, spatial_transformer_sampler
    leaves = [(e["x"].detach().requires_grad_(pb.need_gx), e["theta"].detach().requires_grad_(True)) for e in pb.sets]
    keep = []

    def step(e, xl, tl):
        tp = rotation_dropout(tl, ratio=0.0) if pb.mask01 == 0.0 else tl
        points = spatial_transformer_grid(tp, (pb.oH, pb.oW))
        rois = spatial_transformer_sampler(xl, points)
        torch.autograd.backward([rois], [e["gy"]])
        keep.append((rois, points))                    # every set keeps its own outputs, like the headline's buffers

    n0 = pb._lib.launch_count()
    step(pb.sets[0], *leaves[0])
    launches = pb._lib.launch_count() - n0
    kb = pb._lib.last_kernel()
    torch.cuda.synchronize()
    ok = bool(torch.equal(leaves[0][1].grad, _fresh_gtheta(pb, pb.sets[0])))
    for xl, tl in leaves:
        xl.grad = None
        tl.grad = None
    del keep[:]
    g = hz.capture(lambda: [step(e, xl, tl) for e, (xl, tl) in zip(pb.sets, leaves)])
    for _ in range(3):
        g.replay()
    ms = pb.per_launch(g, steps)
    out = {"us_per_step": ms * 1e3, "value": hz.world * pb.N / (ms * 1e-3), "unit": UNIT,
           "whole_step_frac": (pb.fwd_bytes + pb.bwd_bytes) / (ms * 1e-3) / 1e9 / peak,
           "kernel_launches_per_step": launches, "bwd_kernel": kb, "gtheta_bitwise_equal_to_the_c_abi_step": ok,
           "api": "loans_b200.functions.rotation_dropout -> spatial_transformer_grid -> spatial_transformer_sampler + "
                  "torch.autograd.backward, CUDA-graph replay (reference sheep/sheep_localizer.py:61-63 as written)"}
    del g, keep[:], leaves
    return out


def _fresh_gtheta(pb, e):
    pb.fwd(e)
    pb.bwd(e)
    pb.hz.torch.cuda.synchronize()
    return e["gtheta"].clone()


def floor_probes(hz, pb, steps, peak):
    """What kernels of our launch shape cost before any STN arithmetic (stn_probe.cu), in the same graph harness."""
    torch = hz.torch
    L, _lib = pb.lib, pb._lib
    ctas = max(1, min(4 * torch.cuda.get_device_properties(hz.dev).multi_processor_count, 8 * pb.N))   # 512 at cfg2: the headline kernels' grid
    S = pb.S
    out = {}

    def graph_of(fn, n):
        return hz.capture(lambda: [fn(i) for i in range(n)])

    def st():
        return torch.cuda.current_stream().cuda_stream

    # (1) two empty nodes per step, programmatic dependent launch like the real ones
    g = graph_of(lambda i: (_lib.check(L.loans_stn_probe(0, None, None, 0, 0, ctas, st()), "probe"),
                            _lib.check(L.loans_stn_probe(0, None, None, 0, 0, ctas, st()), "probe")), S)
    for _ in range(3):
        g.replay()
    out["empty_2_nodes_us"] = pb.per_launch(g, steps) * 1e3
    # (2) two nodes of two dependent DRAM round trips + a store each (inputs rotate through the frame buffers)
    outb = torch.empty(ctas * 256, dtype=torch.float32, device=hz.dev)

    def chain(i):
        x = pb.sets[i % S]["x"]
        for _ in range(2):
            _lib.check(L.loans_stn_probe(1, ptr(x), ptr(outb), x.numel() * 4, outb.numel() * 4, ctas, st()), "probe")
    g = graph_of(chain, S)
    for _ in range(3):
        g.replay()
    out["latency_chain_2_nodes_us"] = pb.per_launch(g, steps) * 1e3
    # (3) the same algorithmic bytes as the fused forward + backward, moved by a plain streaming kernel per launch:
    #     forward reads its taps' bytes from x and writes y + grid bytes; backward reads gy + taps and writes gx
    if pb.need_gx:
        fr = pb.fwd_bytes - (pb.N * pb.C * pb.oH * pb.oW * (2 if pb.bf16 else 4) + pb.N * 2 * pb.oH * pb.oW * 4)
        fw = pb.fwd_bytes - fr
        bw = 4 * pb.B * pb.C * pb.H * pb.Wd
        br = pb.bwd_bytes - bw

        def stream_step(i):
            e = pb.sets[i % S]
            xb = e["x"].numel() * 4
            _lib.check(L.loans_stn_probe(2, ptr(e["x"]), ptr(e["gx"]), min(fr, xb) // 16 * 16, fw // 16 * 16, ctas, st()), "probe")
            _lib.check(L.loans_stn_probe(2, ptr(e["x"]), ptr(e["gx"]), min(br, xb) // 16 * 16, bw // 16 * 16, ctas, st()), "probe")
        g = graph_of(stream_step, S)
        for _ in range(3):
            g.replay()
        us = pb.per_launch(g, steps) * 1e3
        out["stream_same_bytes_2_nodes_us"] = us
        out["stream_same_bytes_frac_of_peak"] = (pb.fwd_bytes + pb.bwd_bytes) / (us * 1e-6) / 1e9 / peak

        # (4) the same bytes behind the two dependent round trips every fused kernel has (theta -> tap addresses -> bytes)
        def chain_stream_step(i):
            e = pb.sets[i % S]
            xb = e["x"].numel() * 4
            _lib.check(L.loans_stn_probe(3, ptr(e["x"]), ptr(e["gx"]), min(fr, xb) // 16 * 16, fw // 16 * 16, ctas, st()), "probe")
            _lib.check(L.loans_stn_probe(3, ptr(e["x"]), ptr(e["gx"]), min(br, xb) // 16 * 16, bw // 16 * 16, ctas, st()), "probe")
        g = graph_of(chain_stream_step, S)
        for _ in range(3):
            g.replay()
        us = pb.per_launch(g, steps) * 1e3
        out["chain_then_stream_2_nodes_us"] = us
        out["chain_then_stream_frac_of_peak"] = (pb.fwd_bytes + pb.bwd_bytes) / (us * 1e-6) / 1e9 / peak

        def stream_one(i, rd, wr):
            e = pb.sets[i % S]
            _lib.check(L.loans_stn_probe(2, ptr(e["x"]), ptr(e["gx"]), min(rd, e["x"].numel() * 4) // 16 * 16, wr // 16 * 16, ctas, st()), "probe")
        for key, rd, wr in (("fwd", fr, fw), ("bwd", br, bw)):
            g = graph_of(lambda i, rd=rd, wr=wr: stream_one(i, rd, wr), S)
            for _ in range(3):
                g.replay()
            out["stream_%s_bytes_1_node_us" % key] = pb.per_launch(g, steps) * 1e3
    out["note"] = ("per step of two graph nodes with the fused kernels' launch attributes (%d CTAs x 256 threads): empty kernels; two "
                   "dependent DRAM loads + a store per thread; the step's algorithmic bytes read / written as plain coalesced "
                   "16-byte accesses (what the roofline's denominator costs on this GPU at THIS size); the same bytes behind the two "
                   "dependent round trips (chain_then_stream: an ideal fused step -- no arithmetic, no gather, no scatter)" % ctas)
    return out


def other_configs(hz, args, peak):
    """Every other BASELINE config on this GPU (device-generated inputs, >= 50 timed steps each)."""
    from loans_b200 import workloads as W
    out = {}
    plan = [("cfg1", True), ("cfg2_nogx", False), ("cfg3", True), ("cfg4", True), ("cfg5", True), ("cfg5_nogx", False)]
    for key, need_gx in plan:
        name = key.split("_")[0]
        wl = W.WORKLOADS[name]._replace(rotation_ratio=0.0)
        pb = PathBench(hz, wl, need_gx, seed=4321, min_sets=2 if name in ("cfg3", "cfg4", "cfg5") else 4).build_graphs(per_set=False)
        for _ in range(2):
            pb.g_all.replay()
        steps = 50
        ms = pb.per_launch(pb.g_all, steps, at_least=True)
        ms_f = pb.per_launch(pb.g_fwd, steps, at_least=True)
        ms_b = pb.per_launch(pb.g_bwd, steps, at_least=True)
        r = pb.fractions(ms, ms_f, ms_b, peak)
        r["timed_steps"] = -(-steps // pb.S) * pb.S
        r["config"] = workload_config(wl, need_gx)
        out[key] = r
        pb.free()
    return out


def cfg5_sharded(hz, args, peak):
    """BASELINE configs[4]: the global batch of 1024 frames sharded over the ranks (rank r takes shard_bounds(1024, N, r)),
    fwd+bwd on the shard, and the step's ONE collective -- the mean all-reduce of the localizer's gradients (12.59 M fp32 =
    50.4 MB) -- started on a side stream after the backward and overlapped with the next step's forward."""
    torch, dist = hz.torch, hz.dist
    from loans_b200 import workloads as W
    from loans_b200.parallel import GradientAllReduce, shard_bounds
    wl = W.WORKLOADS["cfg5"]
    lo, hi = shard_bounds(wl.batch, hz.world, hz.rank)
    pb = PathBench(hz, wl, True, seed=9876, batch=hi - lo, min_sets=2)
    n0 = pb._lib.launch_count()
    pb.fwd(pb.sets[0]); kf = pb._lib.last_kernel()
    pb.bwd(pb.sets[0]); kb = pb._lib.last_kernel()
    torch.cuda.synchronize()
    g_f = [hz.capture(lambda e=e: pb.fwd(e)) for e in pb.sets]
    g_b = [hz.capture(lambda e=e: pb.bwd(e)) for e in pb.sets]
    ar = GradientAllReduce([(LOCALIZER_PARAMS,)], hz.dev)
    ar.flat.normal_()
    S = pb.S

    def steps_stn(k):
        for i in range(k):
            g_f[i % S].replay()
            g_b[i % S].replay()

    def steps_full(k):
        for i in range(k):
            g_f[i % S].replay()            # overlaps the previous step's all-reduce
            ar.finish()                    # the previous step's gradients are reduced (at most one collective in flight)
            g_b[i % S].replay()
            ar.start()                     # side stream, after this step's backward
        ar.finish()

    for fn in (steps_stn, steps_full):
        fn(5)
    torch.cuda.synchronize()
    # steps for a timed region of >= ~60 ms (the shard's step is tens of microseconds at 8 GPUs)
    t_probe = hz.timed(lambda: steps_full(20)) / 20
    k = int(max(args.steps, min(20000, math.ceil(60.0 / max(t_probe, 1e-3)))))
    ms_stn_local = hz.timed(lambda: steps_stn(k), reduce=False)
    per_rank = hz.gather(ms_stn_local / k)
    ms_stn = hz.max_over_ranks(ms_stn_local)
    ms_full = hz.timed(lambda: steps_full(k))
    out = {"global_batch": wl.batch, "shard": [lo, hi], "crops_per_gpu": hi - lo, "steps": k,
           "stn_only": {"value": wl.batch * k / (ms_stn * 1e-3), "unit": UNIT, "us_per_step": ms_stn / k * 1e3,
                        "per_rank_us_per_step": {"min": min(per_rank) * 1e3, "median": sorted(per_rank)[len(per_rank) // 2] * 1e3,
                                                 "max": max(per_rank) * 1e3},
                        "whole_step_frac_per_gpu": (pb.fwd_bytes + pb.bwd_bytes) / (ms_stn / k * 1e-3) / 1e9 / peak},
           "with_gradient_allreduce": {"value": wl.batch * k / (ms_full * 1e-3), "unit": UNIT, "us_per_step": ms_full / k * 1e3},
           "kernels": {"fwd": kf, "bwd": kb}, "scaling": "strong (global batch fixed at 1024)",
           "timed_region_ms": {"stn_only": ms_stn, "with_gradient_allreduce": ms_full}}
    if hz.world > 1:
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ar.start(); ar.finish()
        hz.barrier()
        e0.record()
        for _ in range(reps):
            ar.start(); ar.finish()
        e1.record()
        torch.cuda.synchronize()
        us = hz.max_over_ranks(e0.elapsed_time(e1)) / reps * 1e3
        nbytes = LOCALIZER_PARAMS * 4
        out["allreduce"] = {"bytes": nbytes, "us": us, "bus_gbs": 2.0 * (hz.world - 1) / hz.world * nbytes / (us * 1e-6) / 1e9,
                            "reference_bus_gbs_8_ranks": 725.0, "backend": "nccl", "what": "mean all-reduce of one flat fp32 bucket "
                            "of the localizer's 12.59 M gradients on a side stream (loans_b200.parallel.GradientAllReduce), incl. the 1/N scale"}
    else:
        out["allreduce"] = {"bytes": LOCALIZER_PARAMS * 4, "us": 0.0, "note": "one rank: no collective is launched"}
    pb.free()
    return out


def run_ours(args):
    import torch
    from loans_b200 import _lib
    from loans_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = pick_ratio(args, W.WORKLOADS[args.workload])
    need_gx = not args.no_gx
    steps, warm = args.steps, max(args.warmup, 3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, args.cpu_seconds, wl.rotation_ratio)      # before CUDA is touched: plain host work

    hz = Harness(args)
    dev, dist = hz.dev, hz.dist
    L = _lib.lib()
    if args.no_pdl:
        _lib.pdl(False)
    if args.band != "auto":
        _lib.band_backward(args.band == "on")
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"

    pb = PathBench(hz, wl, need_gx, host_set0=True).build_graphs()
    B, K, C, H, Wd, oH, oW, N, S = pb.B, pb.K, pb.C, pb.H, pb.Wd, pb.oH, pb.oW, pb.N, pb.S
    ydt, bf16, dt_code = pb.ydt, pb.bf16, pb.dt_code
    fwd_bytes, bwd_bytes = pb.fwd_bytes, pb.bwd_bytes
    sets, host0, mask01 = pb.sets, pb.host0, pb.mask01
    timed = hz.timed

    warm_up(hz, pb, warm)
    # ---- the timed region: exactly K steps
    ms_total = timed(lambda: pb.run_steps(steps))
    ms_step = ms_total / steps
    value = world * N * steps / (ms_total * 1e-3)
    # ---- per-kernel durations (the same launches, forward-only and backward-only graphs)
    ms_f = pb.per_launch(pb.g_fwd, steps)
    ms_b = pb.per_launch(pb.g_bwd, steps)
    headline_kernels = dict(pb.kernels)

    # ---- the other theta regime, same buffers: general affine (mask 1) if the headline is as-shipped, and vice versa
    variants = {}
    if not args.no_variants:
        main_mask = pb.mask01
        alt_mask = 1.0 if main_mask == 0.0 else 0.0
        main_graphs = (pb.g_all, pb.g_one, pb.g_fwd, pb.g_bwd, pb.kernels)
        pb.mask01 = alt_mask
        pb.build_graphs(per_set=False)
        for _ in range(3):
            pb.g_all.replay()
        r = pb.fractions(pb.per_launch(pb.g_all, steps), pb.per_launch(pb.g_fwd, steps), pb.per_launch(pb.g_bwd, steps), peak)
        r["mask01"] = alt_mask
        variants["general_affine" if alt_mask == 1.0 else "as_shipped_axis_aligned"] = r
        if alt_mask == 1.0 and K == 1:
            # BASELINE configs[1] as literally worded -- grid + sampler on a general affine theta, rotation terms r ~ U(-0.2, 0.2)
            # as SURVEY.md 8(d) draws them (the workload's own generator), no dropout node -- on the same frames
            kept = [e["theta"] for e in sets]
            for i, e in enumerate(sets):
                e["theta"] = torch.from_numpy(_theta_only(W, wl, 555 + i, B, True)).to(dev)
            pb.build_graphs(per_set=False)
            for _ in range(3):
                pb.g_all.replay()
            r = pb.fractions(pb.per_launch(pb.g_all, steps), pb.per_launch(pb.g_fwd, steps), pb.per_launch(pb.g_bwd, steps), peak)
            r["mask01"] = 1.0
            r["theta"] = "rotated: r01, r10 ~ U(-0.2, 0.2)"
            variants["general_affine_rotated"] = r
            for e, t in zip(sets, kept):
                e["theta"] = t
        pb.mask01 = main_mask
        pb.g_all, pb.g_one, pb.g_fwd, pb.g_bwd, pb.kernels = main_graphs
        if K == 1 and not bf16:
            variants["drop_in"] = drop_in_variant(hz, pb, steps, peak)
            variants["drop_in"]["vs_headline"] = variants["drop_in"]["us_per_step"] / (ms_step * 1e3)

    # ---- SURVEY.md section 8(f) rows, timed on the same buffers: prepare_images (rank 1), corner points (rank 2), grayscale
    next_rows = {}
    if not args.no_variants:
        prep_out = torch.empty_like(sets[0]["x"])

        def prep(e):                                   # output into the set's own gx buffer: rotating, so the writes reach DRAM
            _lib.check(L.loans_stn_prepare_images(ptr(e["x"]), 255.0, ptr(e["gx"] if need_gx else prep_out), B, C, H, Wd,
                                                  torch.cuda.current_stream().cuda_stream), "prepare_images")
        if C == 3:
            prep(sets[0])
            g_prep = hz.capture(lambda: [prep(e) for e in sets])
            for _ in range(3):
                g_prep.replay()
            ms_p = pb.per_launch(g_prep, steps)
            nb = 8 * B * C * H * Wd
            next_rows["prepare_images"] = {"us": ms_p * 1e3, "frames_per_s": world * B / (ms_p * 1e-3),
                                           "algorithmic_bytes": nb, "achieved_gbs": nb / (ms_p * 1e-3) / 1e9,
                                           "what": "SheepLocalizer.prepare_images as one kernel (uint8 quantise, RGB->BGR, mean), x*255 folded in"}
        if C == 3:
            # the loader's frame path (8f rank 4): decoded uint8 frames as `-r 512` extracts them (384 x 512) -> LANCZOS resize to
            # this workload's frame size -> float32 NCHW / 255, written into the rotating frame buffers
            from loans_b200.functions import FrameIngest
            fh, fw = (384, 512) if (H, Wd) != (384, 512) else (512, 512)
            raw = [torch.randint(0, 256, (B, fh, fw, 3), dtype=torch.uint8, device=dev) for _ in range(min(S, 4))]
            ing = FrameIngest(B, (fh, fw), (H, Wd), device=dev)
            tgt = [e["gx"] if need_gx else prep_out for e in sets]
            ing(raw[0], out=tgt[0])
            g_ing = hz.capture(lambda: [ing(raw[i % len(raw)], out=tgt[i]) for i in range(S)])
            for _ in range(3):
                g_ing.replay()
            ms_i = pb.per_launch(g_ing, steps)
            nb = B * fh * fw * 3 + 4 * B * C * H * Wd
            next_rows["frame_ingest"] = {"us": ms_i * 1e3, "frames_per_s": world * B / (ms_i * 1e-3), "algorithmic_bytes": nb,
                                         "achieved_gbs": nb / (ms_i * 1e-3) / 1e9, "from": [fh, fw], "to": [H, Wd],
                                         "what": "resize_image(frame, image_size) / 255 of the reference's datasets (PIL LANCZOS, bit-exact) "
                                                 "for the batch on the device: uint8 HWC -> float32 NCHW, two integer kernels"}
            del raw, g_ing
        cor = torch.empty((N, 2, 2, 2), dtype=torch.float32, device=dev)
        gcor = torch.randn((N, 2, 2, 2), dtype=torch.float32, device=dev)

        def step_corners(e):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.loans_stn_crop_fwd_corners(ptr(e["x"]), ptr(e["theta"]), float(mask01), ptr(e["y"]), ptr(cor), N, K, C, H, Wd, oH, oW,
                                                    dt_code, st), "crop_fwd_corners")
            _lib.check(L.loans_stn_crop_bwd_corners(ptr(e["x"]), ptr(e["theta"]), float(mask01), ptr(e["gy"]), ptr(gcor), ptr(e["gtheta"]),
                                                    ptr(e["gx"]), N, K, C, H, Wd, oH, oW, dt_code, st), "crop_bwd_corners")
        step_corners(sets[0])
        g_cor = hz.capture(lambda: [step_corners(e) for e in sets])
        for _ in range(3):
            g_cor.replay()
        ms_c = pb.per_launch(g_cor, steps)
        next_rows["corner_points"] = {"us_per_step": ms_c * 1e3, "value": world * N / (ms_c * 1e-3), "unit": UNIT,
                                      "what": "fwd+bwd with points reduced to the grid's four corners (no dense grid written), "
                                              "corner gradient folded into gtheta"}
        if C == 3:
            yg = torch.empty((N, 1, oH, oW), dtype=ydt, device=dev)
            gyg = torch.randn((N, 1, oH, oW), dtype=torch.float32, device=dev).to(ydt)

            def step_gray(e):
                st = torch.cuda.current_stream().cuda_stream
                _lib.check(L.loans_stn_crop_fwd_ex(ptr(e["x"]), ptr(e["theta"]), float(mask01), ptr(yg), None, ptr(cor), _lib.FLAG_GRAY,
                                                   N, K, C, H, Wd, oH, oW, dt_code, st), "crop_fwd_ex")
                _lib.check(L.loans_stn_crop_bwd_ex(ptr(e["x"]), ptr(e["theta"]), float(mask01), ptr(gyg), None, ptr(gcor), ptr(e["gtheta"]),
                                                   ptr(e["gx"]), None, _lib.FLAG_GRAY, N, K, C, H, Wd, oH, oW, dt_code, st), "crop_bwd_ex")
            step_gray(sets[0])
            g_gray = hz.capture(lambda: [step_gray(e) for e in sets])
            for _ in range(3):
                g_gray.replay()
            ms_g = pb.per_launch(g_gray, steps)
            next_rows["grayscale_corners"] = {"us_per_step": ms_g * 1e3, "value": world * N / (ms_g * 1e-3), "unit": UNIT,
                                              "what": "fwd+bwd with the localizer's grayscale epilogue fused (1-channel crops "
                                                      "and gy) and corner points"}

        if C == 3:
            # conv-ready crops (8f rank 3): bf16, planar vs channels-last padded to four channels, same work otherwise
            ycl = torch.empty((N, oH, oW, 4), dtype=torch.bfloat16, device=dev)
            gycl = torch.randn((N, oH, oW, 4), dtype=torch.float32, device=dev).to(torch.bfloat16)
            ypl = torch.empty((N, C, oH, oW), dtype=torch.bfloat16, device=dev)
            gypl = gycl[..., :3].permute(0, 3, 1, 2).contiguous()

            def step_bf16(e, flags):
                st = torch.cuda.current_stream().cuda_stream
                nh = flags == _lib.FLAG_NHWC4
                _lib.check(L.loans_stn_crop_fwd_ex(ptr(e["x"]), ptr(e["theta"]), float(mask01), ptr(ycl if nh else ypl), ptr(e["grid"]), None,
                                                   flags, N, K, C, H, Wd, oH, oW, _lib.BF16, st), "crop_fwd_ex")
                _lib.check(L.loans_stn_crop_bwd_ex(ptr(e["x"]), ptr(e["theta"]), float(mask01), ptr(gycl if nh else gypl), None, None,
                                                   ptr(e["gtheta"]), ptr(e["gx"]), None, flags, N, K, C, H, Wd, oH, oW, _lib.BF16, st), "crop_bwd_ex")
            res = {}
            for key, flags in (("nchw", 0), ("nhwc4", _lib.FLAG_NHWC4)):
                step_bf16(sets[0], flags)
                g_b = hz.capture(lambda flags=flags: [step_bf16(e, flags) for e in sets])
                for _ in range(3):
                    g_b.replay()
                res[key] = pb.per_launch(g_b, steps) * 1e3
            next_rows["bf16_crops_channels_last"] = {"us_per_step_nchw": res["nchw"], "us_per_step_nhwc4": res["nhwc4"],
                                                     "value": world * N / (res["nhwc4"] * 1e-6), "unit": UNIT,
                                                     "what": "fwd+bwd with bf16 crops and gy: planar (N,C,oH,oW) vs channels-last padded to "
                                                             "four channels (N,oH,oW,4), the layout the assessor's first convolution consumes"}

    # ---- the reference's GPU path on the same inputs: cuDNN's spatial-transformer kernels (what chainer calls on a GPU)
    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cudnn and K == 1 and not bf16 and need_gx:
        try:
            from baseline.cudnn_stn import time_cudnn
            us_c, ver, (y_c, gx_c, gt_c) = time_cudnn(wl, sets, float(mask01), max(1, steps // S), dev)
            pb.fwd(sets[0]); pb.bwd(sets[0])
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

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s%s" % (wl.name, "" if need_gx else "_nogx"), {}).get("bwd_dram_bytes")
    # ---- the DRAM write rate of this GPU, measured here: zero-fill of the rotating gx buffers (plain torch fill kernels, a
    # calibration like MEASURED_PEAKS.json's copy, not part of the path).  The backward is write-dominated (gx is dense)
    write_cal = None
    if need_gx:
        g_fill = hz.capture(lambda: [e["gx"].zero_() for e in sets])
        for _ in range(3):
            g_fill.replay()
        ms_w = pb.per_launch(g_fill, steps)
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
    roofline = {"bound": "hbm", "kernel": "backward launch: %s (forward launch: %s)" % (headline_kernels["bwd"], headline_kernels["fwd"]),
                "achieved": ach_b, "peak": peak, "unit": "GB/s", "frac": ach_b / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes, "avg_launch_us": ms_b * 1e3,
                "write_bound": write_cal,
                "fwd_kernel": {"achieved": ach_f, "frac": ach_f / peak, "algorithmic_bytes_per_launch": fwd_bytes,
                               "avg_launch_us": ms_f * 1e3},
                "whole_step": {"achieved": ach_s, "frac": ach_s / peak, "algorithmic_bytes": fwd_bytes + bwd_bytes,
                               "us": ms_step * 1e3}}
    floor = None
    if not args.no_floor:
        floor = floor_probes(hz, pb, steps, peak)

    # ---- e2e: host-buffer API, pinned host buffers, H2D + D2H inside the timed region, every step
    e2e = None
    if not args.no_e2e:
        from loans_b200.functions import stn_crop
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
        # pinned host buffers from the pipeline: inputs and results of a step are views of one pinned allocation each, so a step
        # is ONE copy per direction (the bytes are the same tensors' bytes)
        outs = [pipe.new_host_outputs() for _ in range(2)]
        hin = pipe.new_host_inputs()
        hin["x"].copy_(hx); hin["theta"].copy_(hth); hin["gy"].copy_(hgy)
        for i in range(4):
            pipe.submit(None, None, None, outs[i % 2], mask01=mask01, inputs=hin)
        pipe.drain()
        hz.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.s_in)
        for i in range(e2e_steps):
            pipe.submit(None, None, None, outs[i % 2], mask01=mask01, inputs=hin)
        e1.record(pipe.s_out)
        pipe.drain()
        torch.cuda.synchronize()
        hz.sampler.window(t0, time.perf_counter())
        ms_e = hz.max_over_ranks(e0.elapsed_time(e1))
        if world > 1:
            dist.barrier()
        ok = bool(torch.equal(outs[(e2e_steps - 1) % 2]["y"], ry))            # pipeline and plain call agree bit for bit
        e2e = {"value": world * N * e2e_steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps, "ms_per_step": ms_e / e2e_steps,
               "api": "loans_b200.pipeline.HostCropPipeline (pinned host tensors in and out; C-ABI fwd+bwd on a compute stream, "
                      "H2D and D2H on their own streams, two buffer sets, one copy per direction and step)",
               "link_gbs_per_direction": {"h2d": pipe.h2d_bytes / (ms_e / e2e_steps * 1e-3) / 1e9,
                                          "d2h": pipe.d2h_bytes / (ms_e / e2e_steps * 1e-3) / 1e9},
               "serial": {"value": world * N * e2e_steps / (ms_serial * 1e-3), "ms_per_step": ms_serial / e2e_steps,
                          "api": "loans_b200.functions.stn_crop + autograd backward, copy-in / run / copy-out one step at a time"},
               "matches_serial_result": ok}
        del pipe
        # what the LoANs step needs (SURVEY.md 8d "no gx"): decoded uint8 frames up (the `/ 255` conversion on the device),
        # frames take no gradient, crops + grid + gtheta down -- same pipeline, same kernels minus gx
        if C == 3 and need_gx and K == 1:
            hu8 = torch.from_numpy((host0["x"] * 255).astype("uint8").transpose(0, 2, 3, 1).copy()).pin_memory()
            pipe2 = HostCropPipeline(B, C, H, Wd, (oH, oW), crops_per_frame=K, need_gx=False, out_dtype=ydt, depth=2, device=dev,
                                     uint8_frames=True)
            outs2 = [pipe2.new_host_outputs() for _ in range(2)]
            hin2 = pipe2.new_host_inputs()
            hin2["x"].copy_(hu8); hin2["theta"].copy_(hth); hin2["gy"].copy_(hgy)
            for i in range(4):
                pipe2.submit(None, None, None, outs2[i % 2], mask01=mask01, inputs=hin2)
            pipe2.drain()
            n2 = e2e_steps * 4
            hz.barrier()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(pipe2.s_in)
            for i in range(n2):
                pipe2.submit(None, None, None, outs2[i % 2], mask01=mask01, inputs=hin2)
            e1.record(pipe2.s_out)
            pipe2.drain()
            torch.cuda.synchronize()
            hz.sampler.window(t0, time.perf_counter())
            ms2 = hz.max_over_ranks(e0.elapsed_time(e1))
            if world > 1:
                dist.barrier()
            e2e["loans_step_uint8_frames_no_gx"] = {
                "value": world * N * n2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / n2, "steps": n2,
                "h2d_bytes_per_step": pipe2.h2d_bytes, "d2h_bytes_per_step": pipe2.d2h_bytes,
                "what": "the same pipeline as LoANs would drive it: decoded uint8 HWC frames uploaded (a quarter of the bytes; `/ 255` "
                        "and the NCHW layout by loans_stn_ingest_u8 on the device), frames take no gradient (no gx computed or "
                        "downloaded), crops + grid + gtheta downloaded.  NOT the headline: the reference arm computes gx"}
            del pipe2
        del outs

    launches_per_step = pb.launches_per_step
    harness = {"l2": "rotating %d distinct input/output sets (%.0f MB each, %.0f MB total > 126 MB L2)"
                     % (S, pb.set_bytes / 1e6, S * pb.set_bytes / 1e6),
               "launch": "CUDA-graph replay of the C-ABI calls loans_stn_crop_fwd + loans_stn_crop_bwd, "
                         + ("plain launches" if args.no_pdl else "programmatic dependent launch (griddepcontrol) between consecutive kernels"),
               "band_backward": args.band, "kernels": headline_kernels,
               "sm_count": torch.cuda.get_device_properties(dev).multi_processor_count}
    pb.free()
    sets = host0 = None                            # (and the locals that alias the sets): the memory goes back before the big configs
    configs = None
    if world == 1 and not args.no_configs:
        configs = other_configs(hz, args, peak)
    cfg5 = None
    if not args.no_cfg5:
        cfg5 = cfg5_sharded(hz, args, peak)

    hz.sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": workload_config(wl, need_gx), "harness": harness,
                "clocks": hz.sampler.summary(), "e2e": e2e, "gpu_launches": launches_per_step * steps,
                "roofline": roofline, "floor": floor, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "variants": variants,
                "configs": configs, "cfg5_sharded": cfg5, "next_rows": next_rows}
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
