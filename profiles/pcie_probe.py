"""What the host link gives: pinned H2D, D2H and both at once, 40 MB transfers (the size class of the e2e path's copies)."""
import torch
dev = torch.device("cuda", 0)
n = 40 * 1024 * 1024 // 4
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, a, b in (("H2D", True, False), ("D2H", False, True), ("both", True, True)):
    run(a, b, 3)
    ms = run(a, b)
    print("%-5s %.3f ms per 40 MB per direction -> %.1f GB/s per direction" % (name, ms, 40 * 1.048576 / ms))
