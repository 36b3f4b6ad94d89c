"""Spike leg of __graft_entry__.smoke(): the committed golden fixture (made by the unmodified reference over the
htslib shim) through the C ABI on cuda:0, compared with the expected SAM / per-target results / SEQ_ERROR count."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "spike_toy")


def smoke_spike(ctx):
    from stochasticsim_b200 import spike as sp
    rd = lambda n: open(os.path.join(GOLD, n), "rb").read()
    hdr, body, names = sp.split_header(rd("in.sam"))
    seqs = sp.parse_fasta(rd("in.fa"))
    targets = sp.parse_spike(rd("in.spike"), names)
    with sp.Spike(ctx, names, seqs) as s:
        out, res, st = s.run_host(body, targets, 434)
        n_se = len(s.seq_errors())
    assert hdr + out == rd("out.sam"), "spiked SAM differs from the reference fixture"
    vcf = rd("truth.vcf")
    assert n_se == vcf.count(b"\tSEQ_ERROR\t"), "SEQ_ERROR record count differs from the reference fixture"
    assert sum(1 for r in res if r.status == sp.T_HIT) == sum(vcf.count(b"\t%s\t" % f) for f in (b"PASS", b"MASKED", b"MASKED_OVL", b"UNDETECTED"))
    stats = rd("stdout.txt").decode()
    assert "alignmentCount (#reads) = %d," % st.alignmentCount in stats and "numberOfLociCovered = %d\n" % st.numberOfLociCovered in stats
    print("[smoke] spike: %d reads, %d covered loci, %d targets hit, %d SEQ_ERROR loci, %d rand() draws; bit-exact vs the reference fixture"
          % (st.n_kept, st.numberOfLociCovered, st.n_hits, n_se, st.rng_draws))
