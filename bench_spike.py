"""Spike workload of bench.py (BASELINE.json configs[1]): chr19-length reference, 150 bp paired
reads at 100x (39.1 M reads, ~14 GB of SAM text), 10,000 SBS spike loci, seed 434, one B200 per rank.

The synthetic input comes from tools/gen_synth.c (counter based, so any coordinate range can be generated on its
own: ranks and host threads build their shards independently).  All compute goes through the C ABI
(ssb_spike_run_device for `value`, ssb_spike_run_host for `e2e`)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
CHR19 = 58_617_616
SPIKE_SEED = 434                       # bin/spikeIn.bash:41


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_int), ("frag_mean", C.c_double), ("frag_sd", C.c_double),
                ("coverage", C.c_double), ("sub_rate", C.c_double), ("indel_rate", C.c_double), ("n_rate", C.c_double),
                ("q0_rate", C.c_double), ("softclip_rate", C.c_double), ("refskip_rate", C.c_double), ("filt_rate", C.c_double),
                ("fa_width", C.c_int), ("lower_frac", C.c_double), ("aux_tags", C.c_int)]


def synth_lib():
    so = os.path.join(ROOT, "tools", "_build", "libsynth.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", ROOT, "tools"], check=True)
    L = C.CDLL(so)
    L.synth_default_params.argtypes = [C.POINTER(SynthParams)]
    L.synth_ref_contig.argtypes = [C.POINTER(SynthParams), C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    L.synth_sam_range.restype = C.c_int64
    L.synth_sam_range.argtypes = [C.POINTER(SynthParams), C.c_int, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.synth_spike_table.restype = C.c_int64
    L.synth_spike_table.argtypes = [C.POINTER(SynthParams), C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                    C.c_double, C.c_double, C.c_void_p, C.c_int64]
    return L


def c2_params(L, seed, coverage):
    p = SynthParams()
    L.synth_default_params(C.byref(p))
    p.seed, p.read_len, p.frag_mean, p.frag_sd, p.coverage = seed, 150, 350.0, 35.0, coverage
    p.sub_rate, p.indel_rate = 0.001, 0.0001
    p.aux_tags = int(os.environ.get("SSB_BENCH_AUX", "0"))          # development only: NM:i / RG:Z fields on every line
    return p


def make_workload(L, seed, contig_len, coverage, n_spikes, name=b"chr19", threads=None, log=None):
    """Returns (ref bytes ndarray, body chunks list[bytes-like ndarray], n_reads, spike text bytes)."""
    p = c2_params(L, seed, coverage)
    ref = np.empty(contig_len + 1, dtype=np.uint8)
    L.synth_ref_contig(C.byref(p), 0, contig_len, 0, contig_len, 0, ref.ctypes.data)
    step = 200_000
    ranges = [(lo, min(contig_len, lo + step)) for lo in range(0, contig_len, step)]
    per_pos = coverage / 150.0 * 420.0 * 1.25 + 64          # generous bytes per reference position

    def gen(rg):
        lo, hi = rg
        cap = int((hi - lo) * per_pos) + (1 << 16)
        buf = np.empty(cap, dtype=np.uint8)
        nr = C.c_int64()
        need = L.synth_sam_range(C.byref(p), 0, name, ref.ctypes.data, contig_len, 0, contig_len, lo, hi, buf.ctypes.data, cap, C.byref(nr))
        if need > cap:
            buf = np.empty(need, dtype=np.uint8)
            L.synth_sam_range(C.byref(p), 0, name, ref.ctypes.data, contig_len, 0, contig_len, lo, hi, buf.ctypes.data, need, C.byref(nr))
        return buf[:need], nr.value

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 8)) as ex:
        parts = list(ex.map(gen, ranges))
    n_reads = sum(nr for _, nr in parts)
    sn = L.synth_spike_table(C.byref(p), 0, name, 0, contig_len, n_spikes, 1, 0.01, 0.5, None, 0)
    sb = C.create_string_buffer(sn + 1)
    L.synth_spike_table(C.byref(p), 0, name, 0, contig_len, n_spikes, 1, 0.01, 0.5, sb, sn)
    if log:
        log(f"[bench] generated {n_reads} reads, {sum(b.size for b, _ in parts)} SAM bytes in {time.perf_counter() - t0:.1f} s")
    return ref[:contig_len], [b for b, _ in parts], n_reads, sb.raw[:sn]


def header_for(name, contig_len):
    return ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n@PG\tID:gen_synth\tPN:gen_synth\n" % (name, contig_len)).encode()


def run(args, D):
    import torch
    import stochasticsim_b200 as ssb
    from stochasticsim_b200 import spike as sp
    from bench import ClockSampler, measured_peaks, log
    peak, peak_src = measured_peaks()
    torch.cuda.set_device(D.local)
    L = synth_lib()
    coverage = 100.0 * args.scale
    contig_len = CHR19
    n_spikes = 10_000
    ref, parts, n_reads, spike_text = make_workload(L, 2 + 1000 * D.rank, contig_len, coverage, n_spikes, log=log)
    n = sum(b.size for b in parts)
    ctx = ssb.Context(D.local)
    ctx.profile_enable(True)
    # pinned host copy (e2e input) and the device-resident copy (`value` input)
    hp = ctx.host_alloc(n + 64)
    off = 0
    for b in parts:
        C.memmove(hp + off, b.ctypes.data, b.size)
        off += b.size
    del parts
    hout = ctx.host_alloc(n + 64)
    d_in = ctx.dev_alloc(n + 64)
    d_out = ctx.dev_alloc(n + 64)
    ctx.h2d(d_in, hp, n)
    ctx.sync()
    names = ["chr19"]
    targets = sp.parse_spike(spike_text, names)
    S = sp.Spike(ctx, names, {"chr19": ref.tobytes()})
    tarr = S.make_targets(targets)
    res = (sp.TargetResult * len(targets))()
    st = sp.Stats()

    def step():
        return S.run_device(d_in, n, d_out, n + 1, tarr, len(targets), SPIKE_SEED, res, st)

    for _ in range(args.warmup):
        out_bytes = step()
    ctx.sync()
    for slot in range(2, 7):
        ctx.profile_read(slot)
    l0 = ctx.launches()
    D.barrier()
    torch.cuda.synchronize()
    stage_ms = {}
    with ClockSampler(D.local) as clk:
        ctx.timer_start()
        for _ in range(args.steps):
            out_bytes = step()
            for k, v in st.as_dict().items():
                if k.startswith("ms_"):
                    stage_ms[k] = stage_ms.get(k, 0.0) + v
        ms = ctx.timer_stop()
    torch.cuda.synchronize()
    D.barrier()
    launches = ctx.launches() - l0
    prof = {name: ctx.profile_read(slot) for name, slot in (("parse", 2), ("emit", 3), ("chain", 4), ("other", 5), ("tally", 6))}
    ms = D.max(ms)
    total_reads = D.sum(float(n_reads))
    stats = st.as_dict()

    # ---- e2e: pinned host SAM body in, spiked SAM body back on the host
    e2e_steps = max(1, min(args.steps, 3))
    outn = C.c_size_t()
    Lb = ssb.lib()

    def e2e_step():
        ssb.check(Lb.ssb_spike_run_host(S.handle, hp, n, hout, n + 1, tarr, len(targets), SPIKE_SEED, res, C.byref(st), C.byref(outn)), ctx.handle)

    e2e_step()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = D.max(time.perf_counter() - t0)
    assert outn.value == out_bytes

    # size-independent properties at full size: every kept read written once, same multiset of bytes up to the spiked bases
    assert stats["alignmentCount"] == stats["n_kept"] and out_bytes == stats["out_bytes"]
    emit_ms, emit_n = prof["emit"]
    parse_ms, parse_n = prof["parse"]
    emit_bytes = 2.0 * out_bytes * args.steps
    res_json = {
        "metric": "sam_reads_spiked_per_s", "value": total_reads * args.steps / (ms / 1e3), "unit": "reads/s",
        "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2 chr19 full-length synthetic 150bp paired reads at %gx with 10k SBS spike loci, per GPU" % coverage
                               + ("" if args.scale == 1.0 else f" (depth scaled x{args.scale})"),
                   "reads_per_gpu": n_reads, "sam_bytes_per_gpu": n, "covered_loci": stats["numberOfLociCovered"], "spike_seed": SPIKE_SEED,
                   "targets_hit": stats["n_hits"], "l2": "input (%.2f GB) larger than L2, no flush" % (n / 1e9), "collective": "none"},
        "e2e": {"value": total_reads * e2e_steps / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(out_bytes),
                "steps": e2e_steps, "api": "ssb_spike_run_host (pinned host SAM body -> spiked SAM body on the host)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "emit_kernel", "achieved": emit_bytes / (emit_ms / 1e3) / 1e9 if emit_ms else None,
                     "peak": peak, "unit": "GB/s", "frac": emit_bytes / (emit_ms / 1e3) / 1e9 / peak if emit_ms else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one emit_kernel launch of this exact workload, from the
                     # ncu --set full capture summarised in profiles/r01_emit_kernel_ncu.txt (15.294 GB + 14.206 GB)
                     "traffic": 29.499585e9 if args.scale == 1.0 else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": 2 * out_bytes, "kernel_ms_avg": emit_ms / max(1, emit_n),
                     "kernel_share_of_step": emit_ms / ms if ms else None,
                     "parse_kernel": {"achieved": n * args.steps / (parse_ms / 1e3) / 1e9 if parse_ms else None, "kernel_ms_avg": parse_ms / max(1, parse_n),
                                      "algorithmic_bytes_per_launch": n},
                     "whole_path": {"algorithmic_bytes_per_step": n + out_bytes, "achieved": (n + out_bytes) * args.steps / (ms / 1e3) / 1e9,
                                    "frac": (n + out_bytes) * args.steps / (ms / 1e3) / 1e9 / peak}},
        "stages_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "kernel_groups_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
        "chain_chunks": stats["chain_mode"],
        "rng_chain": {"ms_per_step": stage_ms.get("ms_chain", 0.0) / args.steps, "draws": stats["rng_draws"],
                      "note": "serial by construction: one glibc rand() stream consumed at every covered locus (stochasticSpike.c:1197)"},
        "clocks": clk.summary(),
    }
    # ---- CPU baseline + parity of the sample (rank 0)
    if D.rank == 0 and not args.no_cpu_baseline:
        res_json["cpu_baseline"] = cpu_baseline(L, S_ctx=(ctx, sp), target_s=15.0)
    S.close()
    for p_ in (hp, hout):
        ctx.host_free(p_)
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)
    ctx.close()
    return res_json


def cpu_exe():
    exe = os.path.join(ROOT, "oracle", "_ref", "stochasticSpike")
    if os.path.exists(exe):
        return exe, "reference", "unmodified stochasticSpike.c -O2 over oracle/shim (htslib I/O + pileup restated, SAM text in)"
    exe = os.path.join(ROOT, "oracle", "_build", "spike_oracle")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return exe, "port", "oracle/spike_oracle.c -O2"


def write_sample(L, td, contig_len, coverage, n_spikes, seed=2):
    ref, parts, n_reads, spike_text = make_workload(L, seed, contig_len, coverage, n_spikes, threads=4)
    with open(os.path.join(td, "s.fa"), "wb") as f:
        f.write(b">chr19\n")
        r = ref.tobytes()
        for i in range(0, len(r), 60):
            f.write(r[i:i + 60] + b"\n")
    with open(os.path.join(td, "s.sam"), "wb") as f:
        f.write(header_for("chr19", contig_len))
        for b in parts:
            f.write(b.tobytes())
    open(os.path.join(td, "s.spike"), "wb").write(spike_text)
    return n_reads


def cpu_run(td):
    exe, kind, what = cpu_exe()
    t0 = time.perf_counter()
    r = subprocess.run([exe, "s.sam", "s.fa", "s.spike", str(SPIKE_SEED), "out.sam"], cwd=td, capture_output=True,
                       env=dict(os.environ, SPIKE_ORACLE_CMDNAME="stochasticSpike"))
    if r.returncode != 0:
        raise RuntimeError("CPU reference failed: " + r.stderr.decode()[:500])
    return time.perf_counter() - t0, kind, what


def cpu_baseline(L, S_ctx=None, target_s=15.0):
    """The reference's CPU code on a bounded sample of the same shape (a sub-region of chr19 at 100x)."""
    with tempfile.TemporaryDirectory() as td:
        n_reads = write_sample(L, td, 40_000, 100.0, 8)
        t, kind, what = cpu_run(td)
        rate = n_reads / t
        region = int(min(3_000_000, max(40_000, rate * target_s / (100.0 / 150.0 / 2 * 2))))
        region = int(min(3_000_000, max(40_000, rate * target_s * 1.5)))         # reads ~ region * 100/150
        n_reads = write_sample(L, td, region, 100.0, max(8, region * 10_000 // CHR19))
        t, kind, what = cpu_run(td)
        out = {"value": n_reads / t, "unit": "reads/s", "cores": 1, "kind": kind, "host_cores_available": os.cpu_count(),
               "sample": f"{n_reads} reads over a {region} bp region at 100x (same generator, same spike density) in {t:.2f} s; {what}; single-threaded as the reference is"}
        if S_ctx is not None:
            ctx, sp = S_ctx
            sam = open(os.path.join(td, "s.sam"), "rb").read()
            hdr, body, names = sp.split_header(sam)
            seqs = sp.parse_fasta(open(os.path.join(td, "s.fa"), "rb").read())
            targets = sp.parse_spike(open(os.path.join(td, "s.spike"), "rb").read(), names)
            with sp.Spike(ctx, names, seqs) as S2:
                got, _, _ = S2.run_host(body, targets, SPIKE_SEED)
            same = (hdr + got) == open(os.path.join(td, "out.sam"), "rb").read()
            out["sample_parity"] = "bit-exact SAM" if same else "MISMATCH"
            assert same, "GPU SAM differs from the CPU reference on the sample"
    return out


def run_reference(args, D):
    if D.rank != 0:
        return None
    L = synth_lib()
    budget = 150.0 / max(1, args.steps + args.warmup)
    with tempfile.TemporaryDirectory() as td:
        n_reads = write_sample(L, td, 40_000, 100.0, 8)
        t, kind, what = cpu_run(td)
        region = int(min(3_000_000, max(40_000, n_reads / t * budget * 1.5)))
        n_reads = write_sample(L, td, region, 100.0, max(8, region * 10_000 // CHR19))
        for _ in range(args.warmup):
            cpu_run(td)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_run(td)
        s = time.perf_counter() - t0
    v = n_reads * args.steps / s
    cb = {"value": v, "unit": "reads/s", "cores": 1, "kind": kind, "host_cores_available": os.cpu_count(),
          "sample": f"{n_reads} reads over a {region} bp region at 100x per step; {what}; single-threaded as the reference is"}
    return {"impl": "reference", "metric": "sam_reads_spiked_per_s", "value": v, "unit": "reads/s", "n_gpus": D.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2 chr19 synthetic 150bp paired reads at 100x with 10k SBS spike loci (bounded sub-region sample)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
