"""GPU parity: libssb200's TNC path (through the C ABI) against the oracle -- bit exact."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import fasta_cases as fc
import oracle_bind as ob
import stochasticsim_b200 as ssb

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    c = ssb.Context(0)
    yield c
    c.close()


def gpu_counts_device(ctx, data: bytes, carry=None, want_carry=False):
    n = len(data)
    d = ctx.dev_alloc(max(n, 1) + 64)
    dc = ctx.dev_alloc(64 * 8)
    ctx.memset(dc, 0, 64 * 8)
    if n:
        buf = C.create_string_buffer(data, n)
        ctx.h2d(d, C.addressof(buf), n)
    cout = ssb.tnc.count_device(ctx, d, n, dc, carry_in=carry, want_carry=want_carry)
    out = np.zeros(64, dtype=np.int64)
    ctx.d2h(out.ctypes.data, dc, 64 * 8)
    ctx.sync()
    ctx.dev_free(d)
    ctx.dev_free(dc)
    return (out, cout) if want_carry else out


@pytest.mark.parametrize("data,expect", fc.KAT)
def test_known_answers(ctx, data, expect):
    got = ssb.tnc.count_host(ctx, data)
    assert np.array_equal(got, ob.tnc_counts(data))
    text = ssb.tnc.format_counts(got)
    assert {l.split("\t")[0]: int(l.split("\t")[1]) for l in text.splitlines() if not l.endswith("\t0")} == expect


def test_golden_fixture(ctx):
    import json
    gold = json.load(open(os.path.join(HERE, "golden", "tnc_golden.json")))
    for case in gold:
        data = bytes.fromhex(case["hex"])
        assert ssb.tnc.format_counts(ssb.tnc.count_host(ctx, data)) == case["stdout"], data
        assert ssb.tnc.format_counts(gpu_counts_device(ctx, data)) == case["stdout"], data


def test_fuzz_small(ctx):
    rng = random.Random(21)
    for i in range(400):
        data = fc.random_fasta(rng, max_lines=40)
        assert np.array_equal(gpu_counts_device(ctx, data), ob.tnc_counts(data)), data


def test_fuzz_cut_anywhere(ctx):
    """A FASTA cut at arbitrary byte offsets, pieces chained through ssb_tnc_carry, gives the same totals
    (this is what host chunking and per-GPU shards rely on)."""
    rng = random.Random(22)
    for i in range(150):
        data = fc.random_fasta(rng, max_lines=30) if i % 3 else fc.genome_like(rng, 3000, width=rng.choice([7, 60, 61]), lower_runs=2)
        want = ob.tnc_counts(data)
        cuts = sorted(rng.randrange(0, len(data) + 1) for _ in range(rng.randint(1, 4))) if data else []
        total = np.zeros(64, dtype=np.int64)
        carry = None
        prev = 0
        for c in cuts + [len(data)]:
            piece = data[prev:c]
            got, carry = gpu_counts_device(ctx, piece, carry=carry, want_carry=True)
            assert carry.as_tuple() == ssb.tnc.carry_after(data[:c]).as_tuple(), (data, c)
            total += got
            prev = c
        assert np.array_equal(total, want), (data, cuts)


def test_device_pieces_scan_ahead_of_fixup(ctx, monkeypatch):
    """A device-resident FASTA larger than one piece: the scans of the pieces run back to back (each takes the three bytes in front of it
    from the buffer), the fix-up kernels follow on a second stream and carry the state from piece to piece, two exception lists alternate.
    Pieces from 32 bytes to 100 kB on inputs with headers, N blocks, lower-case runs and short lines."""
    rng = random.Random(29)
    data = fc.genome_like(rng, 400_000, width=60, n_block=(50_000, 71_003), lower_runs=20, contigs=3)
    want = ob.tnc_counts(data)
    for piece in (4096, 65_536, 100_000):
        monkeypatch.setenv("SSB_TNC_PIECE", str(piece))
        got, carry = gpu_counts_device(ctx, data, want_carry=True)
        assert np.array_equal(got, want), piece
        assert carry.as_tuple() == ssb.tnc.carry_after(data).as_tuple(), piece
    for i in range(60):
        data = fc.random_fasta(rng, max_lines=60)
        monkeypatch.setenv("SSB_TNC_PIECE", str(rng.choice([32, 48, 64, 256])))
        assert np.array_equal(gpu_counts_device(ctx, data), ob.tnc_counts(data)), data


def test_host_chunking(ctx, monkeypatch):
    rng = random.Random(23)
    data = fc.genome_like(rng, 400_000, width=60, n_block=(50_000, 71_003), lower_runs=20, contigs=3)
    want = ob.tnc_counts(data)
    for chunk in (64, 4096, 100_000, 1 << 26):
        monkeypatch.setenv("SSB_TNC_CHUNK", str(chunk))
        assert np.array_equal(ssb.tnc.count_host(ctx, data), want), chunk


def test_many_headers_bed_like(ctx):
    """C3 'restricted' shape: one header + one line per interval (bedtools getfasta)."""
    rng = random.Random(24)
    parts = []
    for i in range(5000):
        parts.append(b">chr1:%d-%d\n" % (i * 1000, i * 1000 + 150))
        parts.append(bytes(rng.choices(b"ACGTN", weights=[10, 10, 10, 10, 1], k=rng.randint(30, 400))) + b"\n")
    data = b"".join(parts)
    assert np.array_equal(gpu_counts_device(ctx, data), ob.tnc_counts(data))


def test_adversarial_exception_overflow(ctx):
    """A header every other byte overflows the exception list; the library must redo the call in safe pieces."""
    data = b">\n" * 700_000 + b"ACGT\n" + b"N\nAC\n" * 300_000
    assert np.array_equal(gpu_counts_device(ctx, data), ob.tnc_counts(data))
    assert np.array_equal(ssb.tnc.count_host(ctx, data), ob.tnc_counts(data))


def test_large_sum_rule_and_linearity(ctx):
    """Size-independent properties at a size the oracle still finishes (64 MB): totals follow the 59/60 rule;
    counts(a ++ b) == counts(a) + counts(b | carry from a)."""
    rng = np.random.default_rng(7)
    lines = 1_000_000
    body = rng.integers(0, 4, size=(lines, 60), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    arr = np.full((lines, 61), ord("\n"), dtype=np.uint8)
    arr[:, :60] = lut[body]
    data = b">chrSyn\n" + arr.tobytes()
    got = ssb.tnc.count_host(ctx, data)
    assert int(got.sum()) == lines * 58 + (lines - 1)
    assert np.array_equal(got, ob.tnc_counts(data))


def test_cli_drop_in(ctx, tmp_path, ref_dir):
    """The C main: same stdout / exit status as the reference binary."""
    exe = os.path.join(ob.ROOT, "stochasticsim_b200", "lib", "tncCountsProfile")
    rng = random.Random(25)
    data = fc.genome_like(rng, 200_000, width=60, n_block=(10_000, 12_345), lower_runs=5, contigs=2)
    p = tmp_path / "t.fa"
    p.write_bytes(data)
    got = subprocess.run([exe, str(p)], capture_output=True)
    assert got.returncode == 0
    assert got.stdout.decode() == ob.tnc_text(ob.tnc_counts(data))
    if ref_dir:
        assert got.stdout == subprocess.run([os.path.join(ref_dir, "tncCountsProfile"), str(p)], capture_output=True).stdout
    assert subprocess.run([exe, str(tmp_path / "missing.fa")], capture_output=True).returncode == 1
    assert subprocess.run([exe], capture_output=True).returncode == 1
