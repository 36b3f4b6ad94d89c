#!/usr/bin/env python
"""Deep-panel shape (BASELINE C5 in small): one window at high depth with dense low-VAF targets; times the GPU path (GPU box only).
usage: python tools/panel_probe.py <window_bp> <depth> <n_targets>"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from stochasticsim_b200 import spike as sp, _lib

def main():
    win, depth, nt = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
    tmp = tempfile.mkdtemp()
    prefix = os.path.join(tmp, "in")
    gen = os.path.join(ROOT, "tools", "_build", "gen_synth")
    t0 = time.time()
    subprocess.run([gen, "out=" + prefix, "seed=5", "contigs=chr1:%d" % (win + 2000), "win=1000:%d" % (win + 1000), "read_len=150", "frag_mean=200", "frag_sd=20",
                    "coverage=%g" % depth, "spikes=%d" % nt, "af=0.005:0.05"], check=True, stderr=subprocess.DEVNULL)
    sam = open(prefix + ".sam", "rb").read()
    hdr, body, names = sp.split_header(sam)
    seqs = sp.parse_fasta(open(prefix + ".fa", "rb").read())
    targets = sp.parse_spike(open(prefix + ".spike", "rb").read(), names)
    print("generated %d bytes, %d targets in %.1f s" % (len(body), len(targets), time.time() - t0), flush=True)
    with _lib.Context(0) as ctx, sp.Spike(ctx, names, seqs) as s:
        for rep in range(2):
            t0 = time.time()
            out, res, st = s.run_host(body, targets, 434)
            d = st.as_dict()
            print("run %d: %.2f s wall; chain_mode %d, reads %d, maxDepth %d, hits %d, draws %d, ms: %s" % (rep, time.time() - t0, st.chain_mode, d["n_kept"], d["maxDepth"], d["n_hits"], st.rng_draws,
                  {k: round(v, 1) for k, v in d.items() if k.startswith("ms_")}), flush=True)

main()
