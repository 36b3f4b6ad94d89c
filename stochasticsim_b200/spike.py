"""Stochastic spike-in: Python face of ssb_spike_* (replaces the pileup loop of stochasticSpike.c:1129-1623).

Only a ctypes binding: parsing of the `.spike` table and of the SAM header mirrors what the C main
(stochasticsim_b200/host/stochasticSpike.c) does, so tests and bench.py can drive the same C ABI.
"""
import ctypes as C
import numpy as np
from ._lib import lib, check

T_HIT, T_NOCOV, T_NOCOV_SILENT, T_TAIL, T_ELSEWHERE = 0, 1, 2, 3, 4
FILTER_NAME = ["NONE", "PASS", "MASKED", "MASKED_OVL", "UNDETECTED"]


class Contig(C.Structure):
    _fields_ = [("name", C.c_char_p), ("len", C.c_int64), ("seq", C.c_void_p)]


class Target(C.Structure):
    _fields_ = [("c_tid", C.c_int32), ("reserved", C.c_int32), ("locus", C.c_int64), ("base", C.c_uint8),
                ("pad", C.c_uint8 * 3), ("af", C.c_float)]


class TargetResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("at_tid", C.c_int32), ("at_pos", C.c_int64), ("locus_index", C.c_int64),
                ("ref_base", C.c_uint8), ("mutant_allele", C.c_uint8), ("filter", C.c_uint8), ("pad", C.c_uint8),
                ("ref_cnt", C.c_int32), ("mut_cnt", C.c_int32), ("err_cnt", C.c_int32 * 4), ("rng_offset", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("alignmentCount", "numberOfLociCovered", "totalFoldCoverage", "maxDepth",
                                          "n_lines", "n_kept", "in_bytes", "out_bytes", "n_runs", "n_hits", "rng_draws", "chain_mode")] + \
               [(n, C.c_float) for n in ("ms_parse", "ms_sort", "ms_emit", "ms_cover", "ms_gather", "ms_rng", "ms_chain",
                                          "ms_patch", "ms_total")] + [("_pad0", C.c_float)] + \
               [(n, C.c_int64) for n in ("rng_k_in", "rng_k_out", "locus_base", "n_forwarded")] + \
               [(n, C.c_float) for n in ("ms_handoff_wait", "ms_phase1", "ms_tally", "ms_exchange")] + \
               [("n_window_retries", C.c_int32), ("_pad1", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("_")}


class Shard(C.Structure):
    """struct ssb_spike_shard: one coordinate range of a cooperative run."""
    _fields_ = [("index", C.c_int32), ("count", C.c_int32), ("lo_tid", C.c_int32), ("hi_tid", C.c_int32),
                ("lo_pos", C.c_int64), ("hi_pos", C.c_int64), ("halo_bytes", C.c_uint64)]


class SeqError(C.Structure):
    _fields_ = [("tid", C.c_int32), ("ref_cnt", C.c_int32), ("pos", C.c_int64), ("locus_index", C.c_int64),
                ("err_cnt", C.c_int32 * 4), ("ref_base", C.c_uint8), ("pad", C.c_uint8 * 7)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    vp, sz = C.c_void_p, C.c_size_t
    L.ssb_spike_create.restype = C.c_int
    L.ssb_spike_create.argtypes = [vp, C.POINTER(Contig), C.c_int, C.POINTER(vp)]
    L.ssb_spike_destroy.restype = None
    L.ssb_spike_destroy.argtypes = [vp]
    for name in ("ssb_spike_run_device", "ssb_spike_run_host"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, sz, vp, sz, C.POINTER(Target), sz, C.c_uint, C.POINTER(TargetResult), C.POINTER(Stats), C.POINTER(sz)]
    for name in ("ssb_spike_run_shard_device", "ssb_spike_run_shard_host"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, C.POINTER(Shard), vp, vp, sz, vp, sz, C.POINTER(Target), sz, C.c_uint, C.POINTER(TargetResult), C.POINTER(Stats), C.POINTER(sz)]
    L.ssb_exchange_local_create.restype = C.c_int
    L.ssb_exchange_local_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.ssb_exchange_nccl_create.restype = C.c_int
    L.ssb_exchange_nccl_create.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.ssb_exchange_destroy.restype = None
    L.ssb_exchange_destroy.argtypes = [vp]
    L.ssb_nccl_unique_id.restype = C.c_int
    L.ssb_nccl_unique_id.argtypes = [vp]
    L.ssb_nccl_comm_init_rank.restype = C.c_int
    L.ssb_nccl_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.ssb_nccl_comm_destroy.restype = None
    L.ssb_nccl_comm_destroy.argtypes = [vp]
    L.ssb_spike_plan_shards.restype = C.c_int
    L.ssb_spike_plan_shards.argtypes = [vp, sz, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int64, C.POINTER(Shard), C.POINTER(sz), C.POINTER(sz)]
    L.ssb_spike_rand.restype = C.c_int
    L.ssb_spike_rand.argtypes = [vp, C.c_uint, C.c_uint64, sz, vp]
    L.ssb_spike_seq_error_count.restype = C.c_int
    L.ssb_spike_seq_error_count.argtypes = [vp, C.POINTER(sz)]
    L.ssb_spike_seq_errors.restype = C.c_int
    L.ssb_spike_seq_errors.argtypes = [vp, C.POINTER(SeqError), sz]
    _bound = True
    return L


def split_header(sam: bytes):
    """(header bytes, body bytes, contig names, sample) as the C main splits them (sam_hdr_read, :971)."""
    p = 0
    while p < len(sam) and sam[p:p + 1] == b"@":
        e = sam.find(b"\n", p)
        p = len(sam) if e < 0 else e + 1
    names = []
    for line in sam[:p].split(b"\n"):
        if line.startswith(b"@SQ\t"):
            for f in line.split(b"\t"):
                if f.startswith(b"SN:"):
                    names.append(f[3:].decode())
    return sam[:p], sam[p:], names


def parse_fasta(fa: bytes):
    """{name: sequence bytes} with white space dropped and case preserved (faidx_fetch_seq64, :215)."""
    out, name, parts = {}, None, []
    for line in fa.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                out[name] = b"".join(parts)
            name, parts = line[1:].split()[0].decode() if line[1:].split() else "", []
        elif name is not None:
            parts.append(bytes(c for c in line if 33 <= c <= 126) if any(not (33 <= c <= 126) for c in line) else line)
    if name is not None:
        out[name] = b"".join(parts)
    return out


def parse_spike(text: bytes, names):
    """getNextTarget (stochasticSpike.c:98-158) over the whole table: [(contig, c_tid, locus, base, af)] in file order."""
    out = []
    for line in text.split(b"\n"):
        if line.startswith(b"#") or not line:
            continue
        tok = [t for t in line.split(b"\t") if t != b""]          # strtok skips empty fields
        if len(tok) < 4:
            continue
        contig = tok[0].decode()
        try:
            locus = int(_atol(tok[1])) - 1
        except ValueError:
            locus = -1
        af = np.float32(_atof(tok[3]))
        out.append((contig, names.index(contig) if contig in names else -1, locus, tok[2][0], float(af)))
    return out


def _atol(b):
    import re
    m = re.match(rb"\s*([+-]?\d+)", b)
    return int(m.group(1)) if m else 0


def _atof(b):
    import re
    m = re.match(rb"\s*([+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|inf(inity)?|nan))", b, re.I)
    return float(m.group(1)) if m else 0.0


class Spike:
    """One ssb_spike: a reference genome (contigs in @SQ order) on one device."""

    def __init__(self, ctx, names, seqs):
        self._L = _bind()
        self.ctx = ctx
        self.names = list(names)
        self._keep = [(n.encode(), np.frombuffer(seqs[n], dtype=np.uint8) if n in seqs else None) for n in names]
        arr = (Contig * max(1, len(names)))()
        for i, (nb, s) in enumerate(self._keep):
            arr[i].name = nb
            arr[i].len = 0 if s is None else s.size
            arr[i].seq = None if s is None or s.size == 0 else s.ctypes.data
        self.handle = C.c_void_p()
        check(self._L.ssb_spike_create(ctx.handle, arr, len(names), C.byref(self.handle)), ctx.handle)

    def close(self):
        if self.handle:
            self._L.ssb_spike_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def make_targets(recs):
        arr = (Target * max(1, len(recs)))()
        for i, (_, c_tid, locus, base, af) in enumerate(recs):
            arr[i].c_tid, arr[i].locus, arr[i].base, arr[i].af = c_tid, locus, base, af
        return arr

    def run_host(self, body: bytes, targets, seed: int):
        """(output body bytes, [TargetResult], Stats) for a host-resident SAM body."""
        n = len(body)
        src = np.frombuffer(body, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
        dst = np.empty(n + 2, dtype=np.uint8)
        tarr = self.make_targets(targets)
        res = (TargetResult * max(1, len(targets)))()
        st, outn = Stats(), C.c_size_t()
        check(self._L.ssb_spike_run_host(self.handle, src.ctypes.data, n, dst.ctypes.data, n + 1, tarr, len(targets),
                                         seed & 0xFFFFFFFF, res, C.byref(st), C.byref(outn)), self.ctx.handle)
        return dst[:outn.value].tobytes(), list(res)[:len(targets)], st

    def run_device(self, d_sam, n, d_out, out_cap, tarr, n_targets, seed, res, st):
        outn = C.c_size_t()
        check(self._L.ssb_spike_run_device(self.handle, d_sam, n, d_out, out_cap, tarr, n_targets, seed & 0xFFFFFFFF,
                                           res, C.byref(st), C.byref(outn)), self.ctx.handle)
        return outn.value

    def run_shard_host(self, shard, xc, body: bytes, targets, seed: int):
        """One shard of a cooperative run (every shard of the group calls this, each from its own thread)."""
        n = len(body)
        src = np.frombuffer(body, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
        dst = np.empty(n + 2, dtype=np.uint8)
        tarr = self.make_targets(targets)
        res = (TargetResult * max(1, len(targets)))()
        st, outn = Stats(), C.c_size_t()
        check(self._L.ssb_spike_run_shard_host(self.handle, C.byref(shard) if shard is not None else None, xc, src.ctypes.data, n, dst.ctypes.data, n + 1,
                                               tarr, len(targets), seed & 0xFFFFFFFF, res, C.byref(st), C.byref(outn)), self.ctx.handle)
        return dst[:outn.value].tobytes(), list(res)[:len(targets)], st

    def run_shard_device(self, shard, xc, d_sam, n, d_out, out_cap, tarr, n_targets, seed, res, st):
        outn = C.c_size_t()
        check(self._L.ssb_spike_run_shard_device(self.handle, C.byref(shard) if shard is not None else None, xc, d_sam, n, d_out, out_cap, tarr, n_targets,
                                                 seed & 0xFFFFFFFF, res, C.byref(st), C.byref(outn)), self.ctx.handle)
        return outn.value

    def rand(self, seed, k0, n):
        out = np.zeros(n, dtype=np.int32)
        check(self._L.ssb_spike_rand(self.handle, seed & 0xFFFFFFFF, k0, n, out.ctypes.data), self.ctx.handle)
        return out

    def seq_errors(self):
        n = C.c_size_t()
        check(self._L.ssb_spike_seq_error_count(self.handle, C.byref(n)))
        arr = (SeqError * max(1, n.value))()
        check(self._L.ssb_spike_seq_errors(self.handle, arr, n.value))
        return list(arr)[:n.value]


def plan_shards(body: bytes, names, count: int, halo_bases: int):
    """[(Shard, body bytes of that shard)] -- ssb_spike_plan_shards over a host-resident SAM body."""
    L = _bind()
    shards = (Shard * count)()
    off = (C.c_size_t * count)()
    ln = (C.c_size_t * count)()
    arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
    buf = np.frombuffer(body, dtype=np.uint8) if len(body) else np.zeros(1, dtype=np.uint8)
    made = L.ssb_spike_plan_shards(buf.ctypes.data, len(body), arr, len(names), count, halo_bases, shards, off, ln)
    if made < 0:
        check(made)
    out = []
    for g in range(made):
        sh = Shard()
        C.memmove(C.byref(sh), C.byref(shards[g]), C.sizeof(Shard))
        out.append((sh, body[off[g]:off[g] + ln[g]]))
    return out


def local_exchange(n: int):
    """n exchange handles for n shards driven by n threads of this process."""
    L = _bind()
    arr = (C.c_void_p * n)()
    check(L.ssb_exchange_local_create(n, arr))
    return [C.c_void_p(arr[i]) for i in range(n)]


def exchange_destroy(xc):
    _bind().ssb_exchange_destroy(xc)


def run_sharded(ctxs, names, seqs, body: bytes, targets, seed: int, count: int, halo_bases: int):
    """Cuts `body` into `count` coordinate shards and runs them as one cooperative group, one thread per shard (ctxs[i % len(ctxs)]
    is the device context of shard i).  Returns (concatenated output, merged [TargetResult], [Stats per shard], [SeqError], n_shards)."""
    import threading
    plan = plan_shards(body, names, count, halo_bases)
    n = len(plan)
    xcs = local_exchange(n)
    outs, errs = [None] * n, [None] * n
    spikes = [Spike(ctxs[g % len(ctxs)], names, seqs) for g in range(n)]

    def work(g):
        try:
            sh, piece = plan[g]
            out, res, st = spikes[g].run_shard_host(sh, xcs[g], piece, targets, seed)
            outs[g] = (out, res, st, spikes[g].seq_errors())
        except Exception as e:  # noqa: BLE001 -- reported to the caller below
            errs[g] = e

    th = [threading.Thread(target=work, args=(g,)) for g in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s_ in spikes:
        s_.close()
    for x in xcs:
        exchange_destroy(x)
    for e in errs:
        if e is not None:
            raise e
    merged = []
    for t in range(len(targets)):
        pick = None
        for g in range(n):
            r = outs[g][1][t]
            if r.status != T_ELSEWHERE:
                assert pick is None, "two shards claim target %d" % t
                pick = r
        assert pick is not None, "no shard claims target %d" % t
        merged.append(pick)
    return b"".join(o[0] for o in outs), merged, [o[2] for o in outs], [e for o in outs for e in o[3]], n
