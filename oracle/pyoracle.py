"""ctypes front-end to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this module; the product package fqzcomp5_b200 never does.

  Codec("oracle")    oracle/liboracle_rans.so      our C restatement (orc_* symbols)
  Codec("ref")       oracle/_ref/libref_rans.so    unmodified reference, as shipped
                                                   (scalar dispatch)
  Codec("ref_simd")  oracle/_ref/libref_rans_simd.so  reference with its AVX2/AVX-512
                                                   kernels dispatched

All three expose the reference's C interface (htscodecs/rANS_static4x16.h:41-50).
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {
    "oracle": (os.path.join(HERE, "liboracle_rans.so"), "orc_"),
    "ref": (os.path.join(HERE, "_ref", "libref_rans.so"), ""),
    "ref_simd": (os.path.join(HERE, "_ref", "libref_rans_simd.so"), ""),
}
REFERENCE_ROOT = "/root/reference"


def build(quiet=True):
    """Compile the oracle (always) and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "-j8", "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def available(kind):
    return os.path.exists(_LIBS[kind][0])


_u8p = C.POINTER(C.c_ubyte)
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


class Codec:
    def __init__(self, kind="oracle"):
        path, pfx = _LIBS[kind]
        if not os.path.exists(path):
            if kind == "oracle" or os.path.isdir(REFERENCE_ROOT):
                build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.kind = kind
        self.lib = L = C.CDLL(path)
        self._bound = getattr(L, pfx + "rans_compress_bound_4x16")
        self._bound.argtypes = [C.c_uint, C.c_int]
        self._bound.restype = C.c_uint
        self._enc = getattr(L, pfx + "rans_compress_to_4x16")
        self._enc.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint), C.c_int]
        self._enc.restype = C.c_void_p
        self._dec = getattr(L, pfx + "rans_uncompress_to_4x16")
        self._dec.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint)]
        self._dec.restype = C.c_void_p

    def bound(self, n, order):
        return int(self._bound(n, order))

    def compress(self, data, order, cap=None):
        """-> compressed bytes, or None where the C call returns NULL."""
        data = bytes(data)
        n = len(data)
        src = C.create_string_buffer(data, max(n, 1))
        if cap is None:
            cap = self.bound(n, order)
        # +1 so the buffer handed over can be given either parity; keep it even-aligned
        dst = C.create_string_buffer(cap + 8)
        sz = C.c_uint(cap)
        r = self._enc(C.addressof(src), n, C.addressof(dst), C.byref(sz), order)
        if not r:
            return None
        return dst.raw[:sz.value]

    def compress_malloc(self, data, order):
        """out == NULL form: the library allocates; we free()."""
        data = bytes(data)
        src = C.create_string_buffer(data, max(len(data), 1))
        sz = C.c_uint(0)
        r = self._enc(C.addressof(src), len(data), None, C.byref(sz), order)
        if not r:
            return None
        out = C.string_at(r, sz.value)
        _libc.free(r)
        return out

    def uncompress(self, comp, ulen=None):
        """ulen=None: library allocates from the stored size; else caller buffer
        of exactly ulen bytes (needed for NOSZ streams)."""
        comp = bytes(comp)
        src = C.create_string_buffer(comp, max(len(comp), 1))
        if ulen is None:
            sz = C.c_uint(0)
            r = self._dec(C.addressof(src), len(comp), None, C.byref(sz))
            if not r:
                return None
            out = C.string_at(r, sz.value)
            _libc.free(r)
            return out
        dst = C.create_string_buffer(max(ulen, 1))
        sz = C.c_uint(ulen)
        r = self._dec(C.addressof(src), len(comp), C.addressof(dst), C.byref(sz))
        if not r:
            return None
        return dst.raw[:sz.value]

    # raw-pointer forms for timing (no Python copies inside the timed region)
    def compress_into(self, src_addr, n, dst_addr, cap, order):
        sz = C.c_uint(cap)
        r = self._enc(src_addr, n, dst_addr, C.byref(sz), order)
        return sz.value if r else -1

    def uncompress_into(self, src_addr, n, dst_addr, ulen):
        sz = C.c_uint(ulen)
        r = self._dec(src_addr, n, dst_addr, C.byref(sz))
        return sz.value if r else -1
