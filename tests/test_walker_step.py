"""CPU: the bit logic of the phase-1 walker step (one-hot first conflict, one-hot next accepted draw, two popcounts) against the
per-locus loop of the reference it replaces.  tests/csrc/walker_step_check.c restates both; see its header."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_walker_step_matches_per_locus_loop(tmp_path):
    exe = str(tmp_path / "walker_step_check")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(HERE, "csrc", "walker_step_check.c")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
