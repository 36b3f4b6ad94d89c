"""Timeline of the streamed host run on the C2 workload (GPU box): SSB_CHAIN_DEBUG host marks per piece."""
import ctypes as C, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench_spike as bs
import stochasticsim_b200 as ssb
from stochasticsim_b200 import spike as sp
L = bs.synth_lib()
cov = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
ref, parts, n_reads, spike_text = bs.make_workload(L, 2, bs.CHR19, cov, 10_000)
n = sum(b.size for b in parts)
ctx = ssb.Context(0)
hp = ctx.host_alloc(n + 64); hout = ctx.host_alloc(n + 64)
off = 0
for b in parts:
    C.memmove(hp + off, b.ctypes.data, b.size); off += b.size
del parts
names = ["chr19"]
targets = sp.parse_spike(spike_text, names)
S = sp.Spike(ctx, names, {"chr19": ref.tobytes()})
tarr = S.make_targets(targets); res = (sp.TargetResult * len(targets))(); st = sp.Stats(); outn = C.c_size_t()
Lb = sp._bind()
def go():
    t0 = time.perf_counter()
    ssb.check(Lb.ssb_spike_run_host(S.handle, hp, n, hout, n + 1, tarr, len(targets), 434, res, C.byref(st), C.byref(outn)), ctx.handle)
    return time.perf_counter() - t0
go()
print("warm run %.1f ms" % (go() * 1e3), flush=True)
os.environ["SSB_CHAIN_DEBUG"] = "2"
print("debug run %.1f ms" % (go() * 1e3), flush=True)
print({k: v for k, v in st.as_dict().items() if k.startswith("ms_")})
