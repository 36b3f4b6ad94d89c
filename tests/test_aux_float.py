"""CPU: which float-typed optional fields (TAG:f:..., TAG:B:f,...) the tokeniser passes through.  htslib parses the value into a 32-bit
float and prints it back with "%g" (SURVEY.md App. D); a line may only be passed through as bytes when that round trip is the identity.
(a) every "%g" image of a float in the normal range must be ACCEPTED (BAM input is printed exactly that way, host/bam_input.h);
(b) everything accepted must BE its own "%g" image."""
import ctypes as C
import random
import struct

import numpy as np

import stochasticsim_b200 as ssb


def _ok(text: bytes) -> bool:
    L = ssb.lib()
    L.ssb_test_aux_ok.restype = C.c_int
    L.ssb_test_aux_ok.argtypes = [C.c_char_p, C.c_size_t]
    return L.ssb_test_aux_ok(text, len(text)) == 0


def _g(x) -> str:
    return "%g" % float(np.float32(x))


def test_percent_g_images_are_accepted():
    rng = random.Random(5)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 100000.0, 999999.0, 1e6, 1e-4, 9.99999e-5, 123456.0, 1234567.0, 0.000123456, 3.14159274, 1e-5, 1e37, -2.5e-37]
    for _ in range(100_000):
        bits = rng.getrandbits(32)
        f = struct.unpack("<f", struct.pack("<I", bits))[0]
        if f != f or abs(f) == float("inf") or (f != 0 and not (1e-37 < abs(f) < 1e37)):
            continue
        vals.append(f)
    for _ in range(50_000):                       # values people actually write: few digits, moderate magnitude
        vals.append(round(rng.uniform(-1000, 1000), rng.randrange(0, 5)))
        vals.append(rng.uniform(0, 1) * 10 ** rng.randrange(-8, 9))
    for v in vals:
        t = _g(v)
        assert _ok(b"de:f:" + t.encode()), (v, t)
    assert _ok(b"XF:B:f," + ",".join(_g(v) for v in vals[:200]).encode())
    for t in (b"inf", b"-inf", b"nan", b"-nan"):
        assert _ok(b"xx:f:" + t)


def test_accepted_text_is_its_own_image():
    rng = random.Random(6)
    digits = "0123456789"
    n_acc = 0
    for _ in range(300_000):
        # random text near the canonical shapes: optional sign, digits, optional point, optional exponent
        t = rng.choice(["", "-"]) + "".join(rng.choice(digits) for _ in range(rng.randrange(1, 8)))
        if rng.random() < 0.6:
            t += "." + "".join(rng.choice(digits) for _ in range(rng.randrange(0, 8)))
        if rng.random() < 0.3:
            t += rng.choice(["e", "E"]) + rng.choice(["+", "-", ""]) + "".join(rng.choice(digits) for _ in range(rng.randrange(1, 4)))
        if _ok(b"de:f:" + t.encode()):
            n_acc += 1
            assert _g(float(t)) == t, t
    assert n_acc > 1000
    for bad in (b"1.0", b"01", b"+1", b"1e5", b"1e+5", b"1e+005", b"0.10", b"1234567", b"0.00001", b"100000e+01", b"1.5E+10", b".5", b"5.", b"", b"-",
                b"1e+38", b"1e-45", b"Inf", b"NaN", b"infinity", b"0x10", b"1,5"):
        assert not _ok(b"de:f:" + bad), bad
    assert not _ok(b"XF:B:f,1.5,2.50")
    assert _ok(b"XF:B:f,1.5,-2e-07,0")
