"""Seeded inputs for the spike path (tools/gen_synth.c) and helpers that run the CPU checkers.

Shared by the oracle tests (CPU) and the parity tests (GPU).  Everything under oracle/ is the CHECKER."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "tools", "_build", "gen_synth")
ORACLE = os.path.join(ROOT, "oracle", "_build", "spike_oracle")
REF = os.path.join(ROOT, "oracle", "_ref", "stochasticSpike")
PRODUCT = os.path.join(ROOT, "stochasticsim_b200", "lib", "stochasticSpike")
# what the GPU path is diffed against: the UNMODIFIED reference over the htslib shim when it has been built (it travels to the
# GPU box as a binary), else the restatement (which tests/test_spike_oracle.py pins to that binary on every input family)
CHECKER = REF if os.path.exists(REF) else ORACLE

# name -> gen_synth arguments (+ optional edits of the .spike table)
CASES = {
    "plain": dict(seed=11, contigs="chr19:30000", coverage=25, read_len=76, frag_mean=180, frag_sd=10, spikes=60, sm="HG1"),
    "overlap_heavy": dict(seed=12, contigs="chr19:20000", coverage=40, read_len=76, frag_mean=95, frag_sd=12, sub=0.02, indel=0.05,
                          nrate=0.004, q0=0.02, softclip=0.1, refskip=0.03, filt=0.08, lower=0.1, spikes=120, alt_mode=1, aux=1),
    "two_contigs_window": dict(seed=13, contigs="chrA:12000,chrB:9000", win="2000:8000", coverage=30, read_len=50, frag_mean=110, frag_sd=15,
                               sub=0.01, indel=0.02, filt=0.05, spikes=50, af="0.2:0.9"),
    "deep_lowvaf": dict(seed=14, contigs="chr1:4000", coverage=600, read_len=100, frag_mean=160, frag_sd=20, sub=0.005, indel=0.01,
                        nrate=0.001, q0=0.005, spikes=40, af="0.005:0.05"),
    "mask_n_ref": dict(seed=15, contigs="chr19:25000", win="5000:15000", mask=1, coverage=25, read_len=76, frag_mean=150, frag_sd=20,
                       sub=0.003, spikes=80, alt_mode=0),
}


def generate(name, outdir, **override):
    """Writes <outdir>/in.{fa,sam,spike}; returns the prefix."""
    args = dict(CASES[name])
    args.update(override)
    prefix = os.path.join(outdir, "in")
    subprocess.run([GEN, "out=" + prefix] + ["%s=%s" % kv for kv in args.items()], check=True, stderr=subprocess.DEVNULL)
    if name == "two_contigs_window":
        # targets the coverage never reaches, the catch-up cascade of stochasticSpike.c:1596-1599, an unknown contig,
        # duplicates, a locus-0 record, a malformed (3-field) line, targets past the end
        lines = open(prefix + ".spike").read().splitlines()
        head = [l for l in lines if l.startswith("#")]
        body = [l for l in lines if not l.startswith("#")]
        a = [l for l in body if l.startswith("chrA")]
        b = [l for l in body if l.startswith("chrB")]
        extra_front = ["chrA\t0\t.\t0.5", "chrA\t500\tG\t0.5", "chrA\t1999\t.\t0.3", "chrA\t2001\tC\t0.5", "chrA\t2002\t.\t0.5",
                       "chrA\t2003\tA\t0.5", "chrA\t2003\tT\t0.4", "chrA\t2004\t.\t1.0", "chrA\t2100\t.\t0", "chrA\t2100\tG"]
        extra_mid = ["chrZ\t100\t.\t0.5", "chrA\t11000\tT\t0.5", "chrA\t3000\t.\t0.5"]
        extra_end = ["chrB\t8990\t.\t0.5", "chrB\t8999\tA\t0.25", "chrQ\t5\t.\t0.1"]
        open(prefix + ".spike", "w").write("\n".join(head + extra_front + a + extra_mid + b + extra_end) + "\n")
    return prefix


def run_cli(exe, prefix, outdir, seed=434, cmdname=None, sam=None):
    """Runs a stochasticSpike-style binary in `outdir`; returns (rc, stdout bytes, out.sam bytes, truth.vcf bytes)."""
    os.makedirs(outdir, exist_ok=True)
    env = dict(os.environ)
    if cmdname:
        env["SPIKE_ORACLE_CMDNAME"] = cmdname
    r = subprocess.run([exe, sam or (prefix + ".sam"), prefix + ".fa", prefix + ".spike", str(seed), "out.sam"], cwd=outdir,
                       capture_output=True, env=env)
    rd = lambda p: open(os.path.join(outdir, p), "rb").read() if os.path.exists(os.path.join(outdir, p)) else None
    return r.returncode, r.stdout, rd("out.sam"), rd("truth.vcf"), r.stderr


def vcf_without_seq_errors(vcf: bytes) -> bytes:
    return b"".join(l + b"\n" for l in vcf.split(b"\n") if l and b"\tSEQ_ERROR\t" not in l)
