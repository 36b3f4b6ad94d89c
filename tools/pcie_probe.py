"""Host link probe (GPU box): pinned H2D alone, D2H alone, both at once on two streams -- is the link full duplex here?"""
import time
import torch
n = 4 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, chunk=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = chunk or n
    for o in range(0, n, step):
        if h2d:
            with torch.cuda.stream(s1):
                d_a[o:o + step].copy_(h_in[o:o + step], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[o:o + step].copy_(d_b[o:o + step], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for name, a, b, c in (("h2d", 1, 0, None), ("d2h", 0, 1, None), ("both", 1, 1, None), ("both 64MiB chunks", 1, 1, 64 << 20), ("h2d 64MiB chunks", 1, 0, 64 << 20)):
    run(a, b, c)
    t = min(run(a, b, c) for _ in range(3))
    print("%-20s %.1f ms  %.1f GB/s per direction" % (name, t * 1e3, n / t / 1e9), flush=True)
