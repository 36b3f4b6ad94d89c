"""ONE spike input cut into coordinate shards: the planner and the in-process exchange on the CPU; on the GPU,
2 / 4 / 7 cooperating shards on one device against the single-shard run and against the reference binary (oracle/_ref, or the
restated oracle when that is absent) -- SAM bytes, truth.vcf, stats block, per-target results, rand() offsets."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

import spike_cases as sc
import stochasticsim_b200 as ssb
from stochasticsim_b200 import spike as sp

CHECKER = sc.CHECKER


def _load(prefix):
    sam = open(prefix + ".sam", "rb").read()
    hdr, body, names = sp.split_header(sam)
    seqs = sp.parse_fasta(open(prefix + ".fa", "rb").read())
    targets = sp.parse_spike(open(prefix + ".spike", "rb").read(), names)
    return hdr, body, names, seqs, targets


# ------------------------------------------------------------------------------------------------ CPU: planner, exchange
def _keys(piece, names):
    out = []
    for line in piece.split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        out.append((names.index(f[2].decode()), int(f[3]) - 1))
    return out


@pytest.mark.parametrize("name,count,halo", [("plain", 2, 200), ("plain", 5, 76), ("two_contigs_window", 4, 120), ("overlap_heavy", 7, 400), ("plain", 64, 0)])
def test_plan_shards_partitions_the_body(name, count, halo, tmp_path):
    prefix = sc.generate(name, str(tmp_path))
    hdr, body, names, seqs, targets = _load(prefix)
    plan = sp.plan_shards(body, names, count, halo)
    assert 1 <= len(plan) <= count
    own = b""
    prev_hi = (0, 0)
    for g, (sh, piece) in enumerate(plan):
        assert sh.index == g and sh.count == len(plan)
        lo = (sh.lo_tid, sh.lo_pos)
        hi = (sh.hi_tid, sh.hi_pos)
        assert lo == prev_hi
        prev_hi = hi
        halo_part, own_part = piece[:sh.halo_bytes], piece[sh.halo_bytes:]
        assert halo_part == b"" or halo_part.endswith(b"\n")
        ko = _keys(own_part, names)
        assert all(lo <= k < hi for k in ko), g                      # own lines start inside the range
        kh = _keys(halo_part, names)
        assert all(k < lo and k[0] == lo[0] and k[1] + halo > lo[1] for k in kh), g      # halo lines: same contig, within `halo` bases
        # the halo is a suffix of everything before the shard's own lines, and holds every line within `halo` bases
        before = body[:len(own)]
        assert before.endswith(halo_part)
        rest = _keys(before[:len(before) - len(halo_part)], names)
        assert not [k for k in rest if k[0] == lo[0] and k[1] + halo > lo[1]], g
        own += own_part
    assert own == body
    assert plan[-1][0].hi_tid == 0x7fffffff


def test_plan_shards_small_and_empty():
    assert len(sp.plan_shards(b"", ["c"], 4, 100)) == 1
    one = b"r1\t0\tc\t5\t60\t4M\t*\t0\t0\tACGT\tIIII\n"
    plan = sp.plan_shards(one, ["c"], 4, 100)
    assert len(plan) == 1 and plan[0][1] == one
    same = one * 6                                                  # six reads with the same key cannot be separated
    assert len(sp.plan_shards(same, ["c"], 3, 100)) == 1


def test_local_exchange_allgather_and_chain():
    n = 5
    xcs = sp.local_exchange(n)
    L = ssb.lib()
    L.ssb_exchange_test_allgather.restype = C.c_int
    L.ssb_exchange_test_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.ssb_exchange_test_relay.restype = C.c_int
    L.ssb_exchange_test_relay.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    got = [None] * n
    relay = [None] * n

    def work(g):
        for rnd in range(3):
            send = np.array([g * 10 + rnd, g], dtype=np.uint64)
            recv = np.zeros(2 * n, dtype=np.uint64)
            assert L.ssb_exchange_test_allgather(xcs[g], send.ctypes.data, recv.ctypes.data, 16) == 0
            got[g] = recv.copy()
            assert list(recv[0::2]) == [h * 10 + rnd for h in range(n)]
        v = C.c_uint64(100 if g == 0 else 0)
        assert L.ssb_exchange_test_relay(xcs[g], C.byref(v)) == 0                # receive from g-1, add 1, send to g+1
        relay[g] = v.value

    th = [threading.Thread(target=work, args=(g,)) for g in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert relay == [100 + g for g in range(n)]
    for x in xcs:
        sp.exchange_destroy(x)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx():
    c = ssb.Context(0)
    yield c
    c.close()


def _res_key(r):
    return (r.status, r.at_tid, r.at_pos, r.locus_index, r.filter, r.ref_cnt, r.mut_cnt, tuple(r.err_cnt), r.rng_offset, r.mutant_allele, r.ref_base)


def _se_key(e):
    return (e.tid, e.pos, e.locus_index, e.ref_cnt, tuple(e.err_cnt), e.ref_base)


SHARD_CASES = [("plain", 2, 200), ("plain", 4, 200), ("overlap_heavy", 2, 600), ("overlap_heavy", 4, 600), ("overlap_heavy", 7, 600),
               ("two_contigs_window", 2, 200), ("two_contigs_window", 4, 200), ("deep_lowvaf", 4, 300), ("mask_n_ref", 4, 200), ("mask_n_ref", 7, 200)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,count,halo", SHARD_CASES)
def test_shards_equal_single_run(name, count, halo, tmp_path, ctx):
    prefix = sc.generate(name, str(tmp_path))
    hdr, body, names, seqs, targets = _load(prefix)
    with sp.Spike(ctx, names, seqs) as s:
        out1, res1, st1 = s.run_host(body, targets, 434)
        se1 = s.seq_errors()
    out, res, sts, se, n = sp.run_sharded([ctx], names, seqs, body, targets, 434, count, halo)
    assert n >= 2
    assert out == out1, "concatenated shard outputs differ from the single run"
    assert [_res_key(r) for r in res] == [_res_key(r) for r in res1]
    assert [_se_key(e) for e in se] == [_se_key(e) for e in se1]
    assert sum(s_.alignmentCount for s_ in sts) == st1.alignmentCount
    assert sum(s_.numberOfLociCovered for s_ in sts) == st1.numberOfLociCovered
    assert sum(s_.totalFoldCoverage for s_ in sts) == st1.totalFoldCoverage
    assert max(s_.maxDepth for s_ in sts) == st1.maxDepth
    # the rand() offset is handed over exactly: every shard starts where its predecessor stopped
    for a, b in zip(sts, sts[1:]):
        assert a.rng_k_out == b.rng_k_in
    assert sts[0].rng_k_in == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,count,chunk,group,slice_", [("plain", 3, 512, 4, 4096), ("mask_n_ref", 4, 256, 3, 64), ("deep_lowvaf", 2, 128, 2, 16),
                                                           ("overlap_heavy", 4, 512, 4, 4096), ("two_contigs_window", 3, 640, 2, 9)])
def test_shards_with_window_maps(name, count, chunk, group, slice_, tmp_path, ctx, monkeypatch):
    """The chunked chain inside every shard: shards behind the first simulate a window of entry offsets around the expected one,
    prepare the entry -> exit table, and the exact offset travels as a lookup per shard."""
    prefix = sc.generate(name, str(tmp_path))
    hdr, body, names, seqs, targets = _load(prefix)
    with sp.Spike(ctx, names, seqs) as s:
        out1, res1, st1 = s.run_host(body, targets, 434)
    monkeypatch.setenv("SSB_CHAIN_CHUNK", str(chunk))
    monkeypatch.setenv("SSB_CHAIN_GROUP", str(group))
    monkeypatch.setenv("SSB_CHAIN_SLICE", str(slice_))
    out, res, sts, se, n = sp.run_sharded([ctx], names, seqs, body, targets, 434, count, 600)
    assert out == out1
    assert [_res_key(r) for r in res] == [_res_key(r) for r in res1]
    if name in ("plain", "mask_n_ref", "deep_lowvaf"):
        assert any(s_.chain_mode > 1 for s_ in sts[1:]), "expected a shard behind the first to use the window maps"


@pytest.mark.gpu
@pytest.mark.parametrize("name,count,chunk,resident", [("plain", 3, 96, 2), ("mask_n_ref", 4, 128, 3), ("deep_lowvaf", 2, 64, 1)])
def test_shards_with_geometry_fitted_to_a_small_device(name, count, chunk, resident, tmp_path, ctx, monkeypatch):
    """A shard behind the first has windows several times as wide as the first one's; when their slices would not all be resident it takes
    longer groups.  A "device" of one to three blocks makes the small inputs take that path in every shard."""
    prefix = sc.generate(name, str(tmp_path))
    hdr, body, names, seqs, targets = _load(prefix)
    with sp.Spike(ctx, names, seqs) as s:
        out1, res1, st1 = s.run_host(body, targets, 434)
    monkeypatch.setenv("SSB_CHAIN_CHUNK", str(chunk))
    monkeypatch.setenv("SSB_P1_RESIDENT", str(resident))
    out, res, sts, se, n = sp.run_sharded([ctx], names, seqs, body, targets, 434, count, 600)
    assert out == out1
    assert [_res_key(r) for r in res] == [_res_key(r) for r in res1]
    assert any(s_.chain_mode > 1 for s_ in sts[1:]), "expected a shard behind the first to use the window maps"


@pytest.mark.gpu
@pytest.mark.parametrize("name,shards", [("plain", 2), ("overlap_heavy", 4), ("two_contigs_window", 4), ("mask_n_ref", 3)])
def test_cli_shards_match_reference_binary(name, shards, tmp_path, ctx):
    """The drop-in main with SSB_SHARDS=N (N logical shards on the devices at hand) against the reference binary: same SAM, same
    truth.vcf (every line), same stats block."""
    prefix = sc.generate(name, str(tmp_path))
    want = sc.run_cli(CHECKER, prefix, str(tmp_path / "ref"), cmdname="stochasticSpike")
    os.environ["SSB_SHARDS"] = str(shards)
    os.environ["SSB_HALO"] = "600"
    try:
        got = sc.run_cli(sc.PRODUCT, prefix, str(tmp_path / "gpu"))
    finally:
        del os.environ["SSB_SHARDS"], os.environ["SSB_HALO"]
    assert want[0] == 0 and got[0] == 0, got[4]
    assert got[2] == want[2], "SAM differs"
    assert got[1] == want[1], "stats differ: %r vs %r" % (want[1], got[1])
    assert got[3] == want[3], "truth.vcf differs"


@pytest.mark.gpu
def test_halo_too_small_is_detected_and_cli_recovers(tmp_path, ctx):
    prefix = sc.generate("plain", str(tmp_path))
    hdr, body, names, seqs, targets = _load(prefix)
    with pytest.raises(ssb.SSBError) as e:
        sp.run_sharded([ctx], names, seqs, body, targets, 434, 3, 10)          # reads are 76 bases long
    assert e.value.code in (-11, -12)
    want = sc.run_cli(CHECKER, prefix, str(tmp_path / "ref"), cmdname="stochasticSpike")
    os.environ["SSB_SHARDS"] = "3"
    os.environ["SSB_HALO"] = "10"
    try:
        got = sc.run_cli(sc.PRODUCT, prefix, str(tmp_path / "gpu"))            # cut again with a larger halo
    finally:
        del os.environ["SSB_SHARDS"], os.environ["SSB_HALO"]
    assert got[0] == 0 and got[2] == want[2] and got[3] == want[3]


@pytest.mark.gpu
def test_two_devices_when_present(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one device")
    prefix = sc.generate("overlap_heavy", str(tmp_path))
    want = sc.run_cli(CHECKER, prefix, str(tmp_path / "ref"), cmdname="stochasticSpike")
    os.environ["SSB_GPUS"] = "2"
    os.environ["SSB_HALO"] = "600"
    try:
        got = sc.run_cli(sc.PRODUCT, prefix, str(tmp_path / "gpu"))
    finally:
        del os.environ["SSB_GPUS"], os.environ["SSB_HALO"]
    assert got[0] == 0 and got[2] == want[2] and got[3] == want[3] and got[1] == want[1]


@pytest.mark.gpu
@pytest.mark.parametrize("name,piece", [("plain", 150_000), ("overlap_heavy", 90_000), ("two_contigs_window", 40_000), ("deep_lowvaf", 300_000), ("mask_n_ref", 64_000),
                                        ("overlap_heavy", 30_000)])
def test_streamed_host_run_equals_whole_run(name, piece, tmp_path, ctx, monkeypatch):
    """ssb_spike_run_host streams a large body through the device in coordinate pieces, each from the exact state its predecessor
    left (no offset windows); here with tiny pieces, so that dozens of hand-overs, halos and forwarded bases happen."""
    prefix = sc.generate(name, str(tmp_path))
    hdr, body, names, seqs, targets = _load(prefix)
    with sp.Spike(ctx, names, seqs) as s:
        monkeypatch.setenv("SSB_NO_STREAM", "1")
        out1, res1, st1 = s.run_host(body, targets, 434)
        se1 = s.seq_errors()
        monkeypatch.delenv("SSB_NO_STREAM")
        monkeypatch.setenv("SSB_STREAM_BYTES", str(piece))
        out2, res2, st2 = s.run_host(body, targets, 434)
        se2 = s.seq_errors()
    assert st2.n_forwarded >= 3, "expected the body to be streamed in several pieces"
    assert out2 == out1
    assert [_res_key(r) for r in res2] == [_res_key(r) for r in res1]
    assert [_se_key(e) for e in se2] == [_se_key(e) for e in se1]
    for f in ("alignmentCount", "numberOfLociCovered", "totalFoldCoverage", "maxDepth", "rng_draws"):
        assert getattr(st2, f) == getattr(st1, f), f


@pytest.mark.gpu
def test_streamed_cli_matches_reference_binary(tmp_path, ctx):
    prefix = sc.generate("overlap_heavy", str(tmp_path))
    want = sc.run_cli(CHECKER, prefix, str(tmp_path / "ref"), cmdname="stochasticSpike")
    os.environ["SSB_STREAM_BYTES"] = "50000"
    try:
        got = sc.run_cli(sc.PRODUCT, prefix, str(tmp_path / "gpu"))
    finally:
        del os.environ["SSB_STREAM_BYTES"]
    assert got[0] == 0, got[4]
    assert got[2] == want[2] and got[1] == want[1] and got[3] == want[3]
