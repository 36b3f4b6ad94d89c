"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and its host-only
entry points agree with the oracle.  No compute call is made here (no GPU in this container)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

import fasta_cases as fc
import oracle_bind as ob
import stochasticsim_b200 as ssb

ROOT = ob.ROOT


def declared_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_loads_and_exports_every_declared_symbol():
    L = C.CDLL(ssb.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_abi_version_and_strerror():
    L = ssb.lib()
    assert L.ssb_abi_version() == 2
    assert b"no CPU fallback" in L.ssb_strerror(-2)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ssb.SSBError) as e:
        ssb.Context(0)
    assert e.value.code == -2


def test_format_matches_oracle():
    rng = np.random.default_rng(3)
    for _ in range(20):
        c = rng.integers(0, 2**40, size=64, dtype=np.int64)
        assert ssb.tnc.format_counts(c) == ob.tnc_text(c)


def test_carry_after_is_a_monoid_action():
    """carry_after(a+b) == carry_after(b, carry_in=carry_after(a)) for any cut."""
    rng = random.Random(9)
    cases = list(fc.EDGE) + [fc.random_fasta(rng) for _ in range(300)]
    for data in cases:
        whole = ssb.tnc.carry_after(data).as_tuple()
        for cut in {0, 1, 2, 3, len(data) // 2, max(0, len(data) - 1), len(data)}:
            if cut > len(data):
                continue
            st = ssb.tnc.carry_after(data[:cut])
            got = ssb.tnc.carry_after(data[cut:], st).as_tuple()
            assert got == whole, (data, cut)


def test_carry_after_meaning():
    # carry = last byte of the nearest kept newline-terminated record; header / N-only records are invisible
    t = ssb.tnc.carry_after(b">c1\nACGT\nNNNN\n>h\n").as_tuple()
    assert t[2] == ord("T") and t[3] == 0
    t = ssb.tnc.carry_after(b"ACGN\nAC").as_tuple()
    assert t[2] == ord("N") and t[3] == 1 and t[4] == ord("A") and t[5] == 1
    t = ssb.tnc.carry_after(b">hdrACGT").as_tuple()
    assert t[2] == 0 and t[3] == 1 and t[4] == ord(">")
