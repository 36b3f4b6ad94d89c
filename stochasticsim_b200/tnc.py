"""Trinucleotide-context scan: Python face of ssb_tnc_* (replaces tncCountsProfile.c:391-483)."""
import ctypes as C
import numpy as np
from ._lib import lib, check, TncCarry

OUT_ORDER = ["ACA", "ACC", "ACG", "ACT", "ATA", "ATC", "ATG", "ATT", "CCA", "CCC", "CCG", "CCT", "CTA", "CTC", "CTG", "CTT",
             "GCA", "GCC", "GCG", "GCT", "GTA", "GTC", "GTG", "GTT", "TCA", "TCC", "TCG", "TCT", "TTA", "TTC", "TTG", "TTT"]


def _buf(data):
    """bytes / bytearray / numpy uint8 -> (address, length, keepalive)."""
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    b = bytes(data)
    keep = C.create_string_buffer(b, len(b)) if len(b) else C.create_string_buffer(1)
    return C.addressof(keep), len(b), keep


def count_host(ctx, data, carry_in=None, want_carry=False):
    """64 context counts (index 16a+4b+c, A<C<G<T) of a host-resident FASTA piece."""
    L = lib()
    addr, n, keep = _buf(data)
    out = np.zeros(64, dtype=np.int64)
    cout = TncCarry()
    check(L.ssb_tnc_count_host(ctx.handle, addr, n, C.byref(carry_in) if carry_in is not None else None,
                               C.byref(cout), out.ctypes.data_as(C.POINTER(C.c_int64))), ctx.handle)
    return (out, cout) if want_carry else out


def count_device(ctx, d_ptr, n, d_counts, carry_in=None, want_carry=False):
    """Adds the windows of the device-resident piece [d_ptr, d_ptr+n) to the 64 int64 device counters."""
    L = lib()
    cout = TncCarry()
    check(L.ssb_tnc_count_device(ctx.handle, d_ptr, n, C.byref(carry_in) if carry_in is not None else None,
                                 C.byref(cout) if want_carry else None, d_counts), ctx.handle)
    return cout if want_carry else None


def carry_after(data, carry_in=None):
    """Scanner state after consuming `data` (host only, counts nothing): used to cut shards."""
    L = lib()
    addr, n, keep = _buf(data)
    out = TncCarry()
    check(L.ssb_tnc_carry_after(addr, n, C.byref(carry_in) if carry_in is not None else None, C.byref(out)))
    return out


def format_counts(counts64):
    """The reference's 32-line stdout (tncCountsProfile.c:452-483)."""
    L = lib()
    c = np.ascontiguousarray(counts64, dtype=np.int64)
    buf = C.create_string_buffer(4096)
    w = L.ssb_tnc_format(c.ctypes.data_as(C.POINTER(C.c_int64)), buf, 4096)
    if w < 0:
        check(w)
    return buf.raw[:w].decode()
