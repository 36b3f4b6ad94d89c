#!/usr/bin/env python
"""Serial chain vs chunked chain on one generated case, for a list of (chunk, group, slice) settings (GPU box only).
usage: python tools/chain_probe.py <case> chunk:group:slice [chunk:group:slice ...]"""
import os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import spike_cases as sc
from stochasticsim_b200 import spike as sp, _lib

def main():
    name = sys.argv[1]
    tmp = tempfile.mkdtemp()
    prefix = sc.generate(name, tmp)
    sam = open(prefix + ".sam", "rb").read()
    hdr, body, names = sp.split_header(sam)
    seqs = sp.parse_fasta(open(prefix + ".fa", "rb").read())
    targets = sp.parse_spike(open(prefix + ".spike", "rb").read(), names)
    key = lambda r: (r.status, r.at_pos, r.filter, r.ref_cnt, r.mut_cnt, tuple(r.err_cnt), r.rng_offset, r.mutant_allele)
    with _lib.Context(0) as ctx, sp.Spike(ctx, names, seqs) as s:
        os.environ["SSB_CHAIN_SERIAL"] = "1"
        out_s, res_s, st_s = s.run_host(body, targets, 434)
        del os.environ["SSB_CHAIN_SERIAL"]
        for spec in sys.argv[2:]:
            c, g, w = spec.split(":")
            os.environ["SSB_CHAIN_CHUNK"], os.environ["SSB_CHAIN_GROUP"], os.environ["SSB_CHAIN_SLICE"] = c, g, w
            sys.stderr.write("---- %s %s\n" % (name, spec)); sys.stderr.flush()
            out_p, res_p, st_p = s.run_host(body, targets, 434)
            bad_t = [i for i, (a, b) in enumerate(zip(res_p, res_s)) if key(a) != key(b)]
            print(name, spec, "mode", st_p.chain_mode, "sam_equal", out_p == out_s, "draws", st_p.rng_draws, st_s.rng_draws, "first bad targets", bad_t[:5], flush=True)

main()
