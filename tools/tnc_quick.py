"""Quick device-resident timing of the TNC scan (development aid; bench.py is the judged harness)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochasticsim_b200 as ssb

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lines = mb * (1 << 20) // 61
rng = np.random.default_rng(1)
arr = np.full((lines, 61), ord("\n"), dtype=np.uint8)
arr[:, :60] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(lines, 60), dtype=np.uint8)]
data = arr.reshape(-1)
n = data.size
with ssb.Context(0) as ctx:
    d = ctx.dev_alloc(n + 64)
    dc = ctx.dev_alloc(512)
    ctx.h2d(d, data.ctypes.data, n)
    ctx.sync()
    for it in range(6):
        ctx.memset(dc, 0, 512)
        ctx.timer_start()
        ssb.tnc.count_device(ctx, d, n, dc)
        ms = ctx.timer_stop()
        print(f"iter {it}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  ({n / ms / 1e6 / 6515.7 * 100:.1f}% of measured HBM copy peak)")
    out = np.zeros(64, dtype=np.int64)
    ctx.d2h(out.ctypes.data, dc, 512)
    ctx.sync()
    print("windows:", int(out.sum()), "expected:", lines * 58 + lines - 1)
