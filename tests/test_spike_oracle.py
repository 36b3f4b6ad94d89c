"""CPU: pins the spike restatement (oracle/spike_oracle.c).

 * against oracle/_ref/stochasticSpike = the UNMODIFIED reference source compiled over oracle/shim/ (only the
   htslib I/O + pileup layer is restated there; "reference logic over shimmed htslib"), byte for byte on SAM,
   truth.vcf and stdout;
 * against the committed golden fixture tests/golden/spike_toy/ (made by tests/golden/make_spike_golden.py with
   that same reference binary), which also travels to the GPU box where /root/reference does not exist;
 * against the glibc rand() known answers of SURVEY.md App. C."""
import os

import numpy as np
import pytest

import oracle_bind as ob
import spike_cases as sc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "spike_toy")


@pytest.mark.parametrize("name", sorted(sc.CASES))
def test_restatement_matches_reference_over_shim(name, tmp_path, ref_dir):
    if ref_dir is None or not os.path.exists(sc.REF):
        pytest.skip("oracle/_ref not built (reference sources not present on this box)")
    prefix = sc.generate(name, str(tmp_path))
    a = sc.run_cli(sc.REF, prefix, str(tmp_path / "ref"))
    b = sc.run_cli(sc.ORACLE, prefix, str(tmp_path / "ora"), cmdname="stochasticSpike")
    assert a[0] == b[0] == 0
    assert a[2] == b[2], "SAM differs"
    assert a[3] == b[3], "truth.vcf differs"
    assert a[1] == b[1], "stdout differs"


def _gold(name):
    return open(os.path.join(GOLD, name), "rb").read()


def test_golden_fixture_bytes(tmp_path):
    import shutil
    for f in ("in.sam", "in.fa", "in.spike"):
        shutil.copy(os.path.join(GOLD, f), tmp_path / f)
    import subprocess
    env = dict(os.environ, SPIKE_ORACLE_CMDNAME="stochasticSpike")
    r = subprocess.run([sc.ORACLE, "in.sam", "in.fa", "in.spike", "434", "out.sam"], cwd=tmp_path, capture_output=True, env=env)
    assert r.returncode == 0
    assert (tmp_path / "out.sam").read_bytes() == _gold("out.sam")
    assert (tmp_path / "truth.vcf").read_bytes() == _gold("truth.vcf")
    assert r.stdout == _gold("stdout.txt")


def test_simple_spike_fixture_has_no_valid_target():
    """toyExample/simple.spike separates POS and ALT with spaces: strtok("\\t") sees 3 fields, every record is skipped."""
    import stochasticsim_b200.spike as sp
    text = _gold("simple.spike")
    assert sp.parse_spike(text, ["chr19"]) == []


def test_glibc_rand_kat():
    r = ob.glibc_rand(434, 0, 5)
    assert list(r) == [43521843, 555289354, 711778510, 1948175914, 749459303]
    assert int(ob.glibc_rand(434, 10**6, 1)[0]) == 1066614699


def test_usage_exit_status():
    import subprocess
    r = subprocess.run([sc.ORACLE], capture_output=True)
    assert r.returncode == 0 and b"Usage" in r.stderr
