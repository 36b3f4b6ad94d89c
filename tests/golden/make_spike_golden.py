"""Regenerates tests/golden/spike_toy/ with the UNMODIFIED reference compiled over oracle/shim/
(oracle/_ref/stochasticSpike, built by `make ref` in the authoring container where /root/reference exists).

    python tests/golden/make_spike_golden.py

Inputs come from tools/gen_synth (seeded); argv is relative so the ##stochasticSpikeCommand= header line does not
depend on where the fixture lives.  simple.spike is the reference's own toyExample/simple.spike (a parser fixture:
it contains no valid record)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "spike_toy")
os.makedirs(OUT, exist_ok=True)
subprocess.run([os.path.join(ROOT, "tools", "_build", "gen_synth"), "out=" + os.path.join(OUT, "in"), "seed=7", "contigs=chr19:2500,chr20:1200",
                "coverage=14", "read_len=60", "frag_mean=100", "frag_sd=14", "sub=0.02", "indel=0.05", "nrate=0.004", "q0=0.02",
                "softclip=0.08", "refskip=0.03", "filt=0.08", "lower=0.1", "spikes=30", "alt_mode=1", "sm=HG00110", "af=0.05:0.9"], check=True)
with open(os.path.join(OUT, "in.spike"), "a") as f:
    f.write("chr20\t1199\t.\t0.5\nchrUn\t5\tA\t0.5\nchr20\t10\tG\n")
r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "stochasticSpike"), "in.sam", "in.fa", "in.spike", "434", "out.sam"], cwd=OUT,
                   capture_output=True, check=True)
open(os.path.join(OUT, "stdout.txt"), "wb").write(r.stdout)
if os.path.exists("/root/reference/toyExample/simple.spike"):
    shutil.copy("/root/reference/toyExample/simple.spike", os.path.join(OUT, "simple.spike"))
print("wrote", sorted(os.listdir(OUT)))
