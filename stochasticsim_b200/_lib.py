"""ctypes binding of include/ssb200.h.  No torch types cross this boundary: plain pointers and sizes."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libssb200.so")


class SSBError(RuntimeError):
    def __init__(self, code, text, detail=""):
        super().__init__(f"ssb200 error {code}: {text}" + (f" [{detail}]" if detail else ""))
        self.code = code


class TncCarry(C.Structure):
    """struct ssb_tnc_carry (include/ssb200.h)."""
    _fields_ = [("started", C.c_uint8), ("prev", C.c_uint8 * 3), ("carry", C.c_uint8),
                ("frag_nonempty", C.c_uint8), ("frag_first", C.c_uint8), ("frag_has_base", C.c_uint8)]

    def as_tuple(self):
        return (self.started, bytes(self.prev), self.carry, self.frag_nonempty, self.frag_first, self.frag_has_base)


_lib = None


def lib():
    """The loaded libssb200.so.  Raises (loudly) when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a). "
                          "stochasticsim_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, sz, i64p, u8p = C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.c_void_p
    sig = {
        "ssb_abi_version": (C.c_int, []),
        "ssb_strerror": (C.c_char_p, [C.c_int]),
        "ssb_last_error": (C.c_char_p, [vp]),
        "ssb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "ssb_ctx_destroy": (None, [vp]),
        "ssb_ctx_device_info": (C.c_int, [vp, C.c_char_p, sz, C.POINTER(C.c_int), C.POINTER(sz)]),
        "ssb_host_alloc": (C.c_int, [vp, sz, C.POINTER(vp)]),
        "ssb_host_free": (None, [vp, vp]),
        "ssb_dev_alloc": (C.c_int, [vp, sz, C.POINTER(vp)]),
        "ssb_dev_free": (None, [vp, vp]),
        "ssb_memcpy_h2d": (C.c_int, [vp, vp, vp, sz]),
        "ssb_memcpy_d2h": (C.c_int, [vp, vp, vp, sz]),
        "ssb_memset_dev": (C.c_int, [vp, vp, C.c_int, sz]),
        "ssb_sync": (C.c_int, [vp]),
        "ssb_timer_start": (C.c_int, [vp]),
        "ssb_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "ssb_kernel_launches": (C.c_uint64, [vp]),
        "ssb_profile_enable": (C.c_int, [vp, C.c_int]),
        "ssb_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
        "ssb_tnc_count_device": (C.c_int, [vp, u8p, sz, C.POINTER(TncCarry), C.POINTER(TncCarry), vp]),
        "ssb_tnc_count_host": (C.c_int, [vp, u8p, sz, C.POINTER(TncCarry), C.POINTER(TncCarry), i64p]),
        "ssb_tnc_carry_after": (C.c_int, [u8p, sz, C.POINTER(TncCarry), C.POINTER(TncCarry)]),
        "ssb_tnc_allreduce": (C.c_int, [vp, vp, vp]),
        "ssb_tnc_format": (C.c_int, [i64p, C.c_char_p, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def check(code, ctx=None):
    if code != 0:
        L = lib()
        detail = L.ssb_last_error(ctx).decode() if ctx else ""
        raise SSBError(code, L.ssb_strerror(code).decode(), detail)


class Context:
    """One ssb_ctx (one CUDA device).  Raises SSBError(-2) when there is no sm_100 device."""

    def __init__(self, device=0):
        self._L = lib()
        self.handle = C.c_void_p()
        check(self._L.ssb_ctx_create(int(device), C.byref(self.handle)))
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._L.ssb_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def info(self):
        name = C.create_string_buffer(128)
        sm, mem = C.c_int(), C.c_size_t()
        check(self._L.ssb_ctx_device_info(self.handle, name, 128, C.byref(sm), C.byref(mem)), self.handle)
        return {"name": name.value.decode(), "sm_count": sm.value, "total_mem": mem.value}

    def host_alloc(self, nbytes):
        p = C.c_void_p()
        check(self._L.ssb_host_alloc(self.handle, nbytes, C.byref(p)), self.handle)
        return p.value

    def host_free(self, p):
        self._L.ssb_host_free(self.handle, p)

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        check(self._L.ssb_dev_alloc(self.handle, nbytes, C.byref(p)), self.handle)
        return p.value

    def dev_free(self, p):
        self._L.ssb_dev_free(self.handle, p)

    def h2d(self, dst, src, nbytes):
        check(self._L.ssb_memcpy_h2d(self.handle, dst, src, nbytes), self.handle)

    def d2h(self, dst, src, nbytes):
        check(self._L.ssb_memcpy_d2h(self.handle, dst, src, nbytes), self.handle)

    def memset(self, dst, value, nbytes):
        check(self._L.ssb_memset_dev(self.handle, dst, value, nbytes), self.handle)

    def sync(self):
        check(self._L.ssb_sync(self.handle), self.handle)

    def timer_start(self):
        check(self._L.ssb_timer_start(self.handle), self.handle)

    def timer_stop(self):
        ms = C.c_float()
        check(self._L.ssb_timer_stop(self.handle, C.byref(ms)), self.handle)
        return ms.value

    def launches(self):
        return int(self._L.ssb_kernel_launches(self.handle))

    def profile_enable(self, on=True):
        check(self._L.ssb_profile_enable(self.handle, 1 if on else 0), self.handle)

    def profile_read(self, slot, reset=True):
        """(total device ms, launches) of one kernel slot (SSB_PROF_* in include/ssb200.h)."""
        ms, n = C.c_double(), C.c_uint64()
        check(self._L.ssb_profile_read(self.handle, slot, C.byref(ms), C.byref(n), 1 if reset else 0), self.handle)
        return ms.value, int(n.value)
