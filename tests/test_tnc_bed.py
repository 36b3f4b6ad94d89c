"""BED-restricted trinucleotide scan (BASELINE config 3): by definition the counts the reference gives on the FASTA that
holds, per interval in BED order, a header line and the interval's bases on one line.  CPU: the .fai-style index.  GPU: the device
path against the reference binary (or the restated oracle) run on that FASTA, for whole ranges and for partitions of the intervals."""
import os
import random
import subprocess

import numpy as np
import pytest

import fasta_cases as fc
import oracle_bind as ob
import stochasticsim_b200 as ssb
from stochasticsim_b200 import tnc

ROOT = ob.ROOT
REF = os.path.join(ROOT, "oracle", "_ref", "tncCountsProfile")


def _contig_seqs(data):
    out, name, parts = [], None, []
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                out.append((name, b"".join(parts)))
            f = line[1:].split()
            name, parts = (f[0].decode() if f else ""), []
        elif name is not None:
            parts.append(line.rstrip(b"\r"))
    if name is not None:
        out.append((name, b"".join(parts)))
    return out


def virtual_fasta(data, intervals):
    seqs = _contig_seqs(data)
    return b"".join(b">%s:%d-%d\n" % (seqs[c][0].encode(), a, b) + seqs[c][1][a:b] + b"\n" for c, a, b in intervals)


def random_intervals(rng, seqs, n, lo=1, hi=300):
    out = []
    for _ in range(n):
        c = rng.randrange(len(seqs))
        L = len(seqs[c][1])
        if L < 2:
            continue
        ln = min(L, rng.randint(lo, max(lo, hi)))
        a = rng.randrange(0, L - ln + 1)
        out.append((c, a, a + ln))
    return sorted(out)


def test_fasta_index_matches_the_text():
    rng = random.Random(11)
    for k in range(60):
        data = fc.genome_like(rng, rng.choice([300, 5000, 30000]), width=rng.choice([60, 61, 70, 9]), n_block=(50, 90), lower_runs=3, contigs=rng.choice([1, 3, 6]))
        if k % 4 == 0:
            data = data.rstrip(b"\n")
        seqs = _contig_seqs(data)
        idx = tnc.fasta_index(data)
        assert [(n, c.len) for n, c in idx] == [(n, len(s)) for n, s in seqs]
        for (n, c), (_, s) in zip(idx, seqs):
            for p in {0, c.len // 3, c.len - 1} if c.len else set():
                b = c.seq_off + p + (p // c.line_bases) * (c.line_bytes - c.line_bases)
                assert data[b:b + 1] == s[p:p + 1]


@pytest.fixture(scope="module")
def ctx():
    c = ssb.Context(0)
    yield c
    c.close()


def _gpu_counts(ctx, data, idx, intervals, ranges):
    import torch
    d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    total = np.zeros(64, dtype=np.int64)
    for a, b in ranges:
        cnt = torch.zeros(64, dtype=torch.int64, device="cuda")
        tnc.count_bed_device(ctx, d.data_ptr(), d.numel(), idx, intervals, cnt.data_ptr(), a, b)
        ctx.sync()
        total += cnt.cpu().numpy()
    return total


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_bed_counts_match_reference_on_the_extracted_fasta(seed, ctx, tmp_path):
    rng = random.Random(100 + seed)
    data = fc.genome_like(rng, rng.choice([4000, 40000, 200000]), width=rng.choice([60, 70]), n_block=(300, 900), lower_runs=8, contigs=rng.choice([1, 3, 5]))
    seqs = _contig_seqs(data)
    iv = random_intervals(rng, seqs, rng.choice([1, 7, 200, 1500]), lo=rng.choice([1, 30]), hi=rng.choice([3, 400, 5000]))
    if seed % 2:                                        # BED order need not be sorted, and intervals may repeat or overlap
        rng.shuffle(iv)
    idx = tnc.fasta_index(data)
    vf = virtual_fasta(data, iv)
    want = ob.tnc_counts(vf)
    if os.path.exists(REF):                             # the unmodified reference on the extracted FASTA
        p = tmp_path / "v.fa"
        p.write_bytes(vf)
        out = subprocess.run([REF, str(p)], capture_output=True, check=True).stdout.decode()
        assert out == ob.tnc_text(want)
    n = len(iv)
    got = _gpu_counts(ctx, data, idx, iv, [(0, n)])
    assert np.array_equal(got, want)
    cuts = sorted({0, n} | {rng.randrange(0, n + 1) for _ in range(3)})
    parts = _gpu_counts(ctx, data, idx, iv, list(zip(cuts, cuts[1:])))
    assert np.array_equal(parts, want), "partition of the intervals does not add up"


@pytest.mark.gpu
def test_bed_cli(ctx, tmp_path):
    rng = random.Random(5)
    data = fc.genome_like(rng, 60000, width=60, n_block=(500, 1500), lower_runs=6, contigs=3)
    seqs = _contig_seqs(data)
    iv = random_intervals(rng, seqs, 300, lo=30, hi=600)
    (tmp_path / "g.fa").write_bytes(data)
    (tmp_path / "t.bed").write_text("".join("%s\t%d\t%d\n" % (seqs[c][0], a, b) for c, a, b in iv))
    exe = os.path.join(ROOT, "stochasticsim_b200", "lib", "tncCountsProfile")
    got = subprocess.run([exe, str(tmp_path / "g.fa"), str(tmp_path / "t.bed")], capture_output=True)
    assert got.returncode == 0, got.stderr
    assert got.stdout.decode() == ob.tnc_text(ob.tnc_counts(virtual_fasta(data, iv)))
    bad = subprocess.run([exe, str(tmp_path / "g.fa"), str(tmp_path / "t.bed")], capture_output=True, env=dict(os.environ, SSB_DEVICE="99"))
    assert bad.returncode == 3
