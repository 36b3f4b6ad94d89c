"""Regenerates the golden fixtures by RUNNING THE REFERENCE (oracle/_ref/, compiled from
/root/reference by `make ref`).  Only runnable in the authoring container; the fixtures are committed.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fasta_cases as fc  # noqa: E402


def tnc():
    exe = os.path.join(ROOT, "oracle", "_ref", "tncCountsProfile")
    rng = random.Random(2024)
    cases = [d for d, _ in fc.KAT] + fc.EDGE + [fc.random_fasta(rng) for _ in range(60)]
    out = []
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "f.fa")
        for data in cases:
            open(p, "wb").write(data)
            out.append({"hex": data.hex(), "stdout": subprocess.run([exe, p], capture_output=True, check=True).stdout.decode()})
    json.dump(out, open(os.path.join(HERE, "tnc_golden.json"), "w"), indent=0)
    print("tnc_golden.json:", len(out), "cases")


if __name__ == "__main__":
    tnc()
