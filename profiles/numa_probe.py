"""Host topology of the GPU box and what it does to pinned-memory copies: NUMA nodes, the GPUs' nodes, and H2D / D2H / both
at once per GPU with the process (a) left alone, (b) bound (CPU affinity + MPOL_BIND) to each NUMA node in turn.
With --all every visible GPU runs the copy loop at the same time (one process per GPU) so the aggregate ceiling shows.
usage: numa_probe.py [--all] [--bind auto|none|<node>]"""
import ctypes
import glob
import json
import multiprocessing as mp
import os
import subprocess
import sys


def nodes():
    out = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(d.rsplit("node", 1)[1])
        cpus = open(d + "/cpulist").read().strip()
        out[n] = cpus
    return out


def parse_cpulist(s):
    r = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            r += list(range(int(a), int(b) + 1))
        elif part:
            r.append(int(part))
    return r


def bind(node, cpulist):
    """CPU affinity to the node's CPUs (those this process may use) + MPOL_BIND of future allocations to the node."""
    allowed = os.sched_getaffinity(0)
    want = set(parse_cpulist(cpulist)) & allowed
    if want:
        os.sched_setaffinity(0, want)
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(238, 2, ctypes.byref(mask), 64)                  # set_mempolicy(MPOL_BIND, &mask, maxnode)
    return {"cpus": len(want), "set_mempolicy_rc": rc, "errno": ctypes.get_errno() if rc else 0}


def gpu_node(idx):
    try:
        import torch
        bus = torch.cuda.get_device_properties(idx)
        pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
        return pci, int(open("/sys/bus/pci/devices/%s/numa_node" % pci).read())
    except Exception as e:                                              # noqa: BLE001
        return str(e), -1


def copy_rates(idx, barrier=None, reps=20, mb=40):
    import torch
    torch.cuda.set_device(idx)
    n = mb * 1024 * 1024 // 4
    h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    h_in.fill_(1.0)
    h_out.fill_(0.0)
    d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        for timed in (False, True):
            torch.cuda.synchronize()
            if barrier is not None and timed:
                barrier.wait()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_stream(torch.cuda.current_stream())
            s2.wait_stream(torch.cuda.current_stream())
            for _ in range(reps if timed else 3):
                if a:
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if b:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            if timed:
                res[name] = round(mb * 1.048576 / (e0.elapsed_time(e1) / reps), 1)           # GB/s per direction
    return res


def worker(idx, mode, barrier, q):
    nd = nodes()
    info = {"gpu": idx}
    pci, gn = gpu_node(idx)
    info["pci"], info["gpu_numa_node"] = pci, gn
    if mode == "auto" and gn >= 0 and gn in nd:
        info["bound_to"] = gn
        info["bind"] = bind(gn, nd[gn])
    elif mode not in ("auto", "none"):
        info["bound_to"] = int(mode)
        info["bind"] = bind(int(mode), nd[int(mode)])
    info["gbs_per_direction"] = copy_rates(idx, barrier)
    q.put(info)


def main():
    nd = nodes()
    print(json.dumps({"numa_nodes": nd, "cpus_allowed": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}))
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=60).stdout
            print("\n".join(l for l in txt.splitlines() if cmd[0] != "lscpu" or any(k in l for k in ("NUMA", "Socket", "Model name", "CPU(s):", "Thread"))))
        except Exception as e:                                          # noqa: BLE001
            print(cmd, "failed:", e)
    import torch
    ngpu = torch.cuda.device_count()
    all_gpus = "--all" in sys.argv
    modes = ["none"] + (["auto"] if len(nd) > 1 else []) + ([str(n) for n in nd] if len(nd) > 1 and not all_gpus else [])
    if "--bind" in sys.argv:
        modes = [sys.argv[sys.argv.index("--bind") + 1]]
    ctx = mp.get_context("spawn")
    for mode in modes:
        gpus = list(range(ngpu)) if all_gpus else [0]
        barrier = ctx.Barrier(len(gpus)) if len(gpus) > 1 else None
        q = ctx.Queue()
        ps = [ctx.Process(target=worker, args=(i, mode, barrier, q)) for i in gpus]
        for p in ps:
            p.start()
        rows = sorted((q.get(timeout=300) for _ in ps), key=lambda r: r["gpu"])
        for p in ps:
            p.join()
        agg = {k: round(sum(r["gbs_per_direction"][k] for r in rows), 1) for k in ("h2d", "d2h", "both")}
        print(json.dumps({"bind": mode, "gpus": len(gpus), "aggregate_gbs_per_direction": agg, "per_gpu": rows}))


if __name__ == "__main__":
    main()
