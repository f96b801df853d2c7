"""ctypes binding of libb200rans.so (include/b200rans.h).

The function names and argument meaning mirror the reference's C interface
(htscodecs/rANS_static4x16.h:41-64), so the parity tests read like calls to the
reference.  There is no Python or CPU implementation behind these calls: if the
CUDA library is missing or no GPU is usable they raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200rans.so")

RANS_ORDER_PACK = 0x80
RANS_ORDER_RLE = 0x40
RANS_ORDER_CAT = 0x20
RANS_ORDER_NOSZ = 0x10
RANS_ORDER_STRIPE = 0x08
RANS_ORDER_X32 = 0x04
RANS_ORDER_STRIPE_NO0 = 1 << 16
RANS_ORDER_SIMD_AUTO = 1 << 17

EXPORTS = [
    "rans_compress_bound_4x16", "rans_compress_to_4x16", "rans_compress_4x16",
    "rans_uncompress_to_4x16", "rans_uncompress_4x16", "rans_set_cpu",
    "b200rans_set_device", "b200rans_device_count", "b200rans_host_alloc", "b200rans_host_free",
    "b200rans_compress_batch", "b200rans_uncompress_batch", "b200rans_uncompressed_size",
    "b200rans_compress_batch_dev_bound", "b200rans_compress_batch_dev",
    "b200rans_uncompress_batch_dev", "b200rans_compress_batch_multi",
    "b200rans_uncompress_batch_multi", "b200rans_launch_count", "b200rans_version",
    "b200rans_set_profiling", "b200rans_last_kernel_ms", "b200rans_dec_staged_stats",
    "b200rans_compress_methods_batch", "b200rans_compress_methods", "b200rans_compress_trials",
    "b200fq_split", "b200fq_join", "b200fq_split_dev", "b200fq_join_dev",
    "b200fq_split_scratch_bytes", "b200fq_join_scratch_bytes",
    "b200fqz_crc32", "b200fqz_crc32_dev", "b200fqz_assemble_block_dev",
    "b200rans_compress_slots_bound", "b200rans_compress_batch_dev2", "b200rans_compress_trials_dev",
    "b200rans_tok3_methods",
    "b200fq_split_mode", "b200fq_split_dev_mode",
    "b200fqz_block_bound", "b200fqz_encode_block", "b200fqz_decode_block",
    "b200fqz_encode_blocks_multi", "b200fqz_decode_blocks_multi",
    "b200fqz_learner_init", "b200fqz_learner_methods", "b200fqz_learner_update",
]

_lib = None
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


class B200RansError(RuntimeError):
    pass


def lib():
    """Load the CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200RansError(
                "libb200rans.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "fqzcomp5_b200 has no CPU path")
        L = C.CDLL(LIB_PATH)
        vp, u32, i32, sz = C.c_void_p, C.c_uint, C.c_int, C.c_size_t
        pu32, pi32 = C.POINTER(C.c_uint), C.POINTER(C.c_int)
        L.rans_compress_bound_4x16.argtypes = [u32, i32]
        L.rans_compress_bound_4x16.restype = u32
        L.rans_compress_to_4x16.argtypes = [vp, u32, vp, pu32, i32]
        L.rans_compress_to_4x16.restype = vp
        L.rans_compress_4x16.argtypes = [vp, u32, pu32, i32]
        L.rans_compress_4x16.restype = vp
        L.rans_uncompress_to_4x16.argtypes = [vp, u32, vp, pu32]
        L.rans_uncompress_to_4x16.restype = vp
        L.rans_uncompress_4x16.argtypes = [vp, u32, pu32]
        L.rans_uncompress_4x16.restype = vp
        L.rans_set_cpu.argtypes = [i32]
        L.b200rans_set_device.argtypes = [i32]
        L.b200rans_host_alloc.argtypes = [sz]
        L.b200rans_host_alloc.restype = vp
        L.b200rans_host_free.argtypes = [vp]
        L.b200rans_compress_batch.argtypes = [i32, vp, vp, vp, vp, sz, vp, vp]
        L.b200rans_uncompress_batch.argtypes = [i32, vp, vp, vp, vp, vp]
        L.b200rans_uncompressed_size.argtypes = [vp, u32]
        L.b200rans_uncompressed_size.restype = C.c_int64
        L.b200rans_compress_batch_dev_bound.argtypes = [i32, vp, vp]
        L.b200rans_compress_batch_dev_bound.restype = sz
        L.b200rans_compress_batch_dev.argtypes = [vp, i32, vp, vp, vp, vp, vp, sz, vp, vp]
        L.b200rans_uncompress_batch_dev.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.b200rans_compress_batch_multi.argtypes = [i32, i32, vp, vp, vp, vp, vp, sz, vp, vp]
        L.b200rans_uncompress_batch_multi.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
        L.b200rans_compress_methods_batch.argtypes = [i32, vp, vp, i32, vp, vp, sz, vp, vp, vp, vp]
        L.b200rans_compress_trials.argtypes = [i32, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp]
        L.b200rans_compress_methods.argtypes = [vp, u32, i32, vp, pu32, pi32, vp]
        L.b200rans_compress_methods.restype = vp
        L.b200fq_split.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp, vp, u32, vp]
        L.b200fq_split_mode.argtypes = [i32, u32, vp, u32, vp, u32, vp, vp, u32, vp, vp, u32, vp]
        L.b200fq_split_dev_mode.argtypes = [vp, i32, u32, vp, u32, vp, u32, vp, vp, u32, vp, vp, vp, vp, u32, vp, sz, vp]
        L.b200fq_join.argtypes = [vp, u32, vp, vp, u32, vp, u32, i32, vp, u32, vp]
        L.b200fq_split_scratch_bytes.argtypes = [u32, u32]
        L.b200fq_split_scratch_bytes.restype = sz
        L.b200fq_join_scratch_bytes.argtypes = [u32, u32]
        L.b200fq_join_scratch_bytes.restype = sz
        L.b200fq_split_dev.argtypes = [vp, vp, u32, vp, u32, vp, vp, u32, vp, vp, vp, vp, u32, vp, sz, vp]
        L.b200fq_join_dev.argtypes = [vp, vp, u32, vp, vp, vp, u32, i32, vp, u32, vp, sz, vp]
        L.b200fqz_crc32.argtypes = [u32, vp, C.c_uint64, pu32]
        L.b200fqz_crc32_dev.argtypes = [vp, vp, C.c_uint64, u32, vp]
        L.b200fqz_assemble_block_dev.argtypes = [vp, u32, i32, vp, vp, C.c_uint64, pu32]
        L.b200rans_compress_slots_bound.argtypes = [i32, vp, vp]
        L.b200rans_compress_slots_bound.restype = sz
        L.b200rans_compress_batch_dev2.argtypes = [vp, i32, vp, vp, vp, vp, vp, sz, vp, vp, u32]
        L.b200rans_compress_trials_dev.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, sz, u32, vp, vp, vp, vp]
        L.b200rans_tok3_methods.argtypes = [i32, i32, u32, vp]
        L.b200fqz_block_bound.argtypes = [u32]
        L.b200fqz_block_bound.restype = sz
        L.b200fqz_encode_block.argtypes = [vp, u32, vp, vp, sz, vp]
        L.b200fqz_decode_block.argtypes = [vp, u32, i32, vp, sz, vp]
        L.b200fqz_encode_blocks_multi.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
        L.b200fqz_decode_blocks_multi.argtypes = [i32, i32, vp, vp, i32, vp, vp, vp]
        L.b200fqz_learner_init.argtypes = [vp]
        L.b200fqz_learner_methods.argtypes = [vp, vp, vp]
        L.b200fqz_learner_update.argtypes = [vp, vp, vp]
        for f in (L.b200fqz_learner_init, L.b200fqz_learner_methods, L.b200fqz_learner_update):
            f.restype = None
        L.b200rans_launch_count.restype = C.c_uint64
        L.b200rans_version.restype = C.c_char_p
        L.b200rans_set_profiling.argtypes = [i32]
        L.b200rans_last_kernel_ms.argtypes = [i32]
        L.b200rans_last_kernel_ms.restype = C.c_float
        L.b200rans_dec_staged_stats.argtypes = [C.POINTER(C.c_ulonglong), i32]
        _lib = L
    return _lib


def _check(rc, what):
    if rc:
        raise B200RansError("%s failed with status %d (see stderr)" % (what, rc))


# ---------------------------------------------------------------- reference-shaped calls
def rans_compress_bound_4x16(size, order):
    return int(lib().rans_compress_bound_4x16(size, order))


def rans_compress_to_4x16(data, order, cap=None):
    """Compress one buffer.  Returns bytes, or None where the C call returns NULL."""
    L = lib()
    data = bytes(data)
    n = len(data)
    src = C.create_string_buffer(data, max(n, 1))
    if cap is None:
        cap = rans_compress_bound_4x16(n, order)
    dst = C.create_string_buffer(cap + 8)
    sz = C.c_uint(cap)
    r = L.rans_compress_to_4x16(C.addressof(src), n, C.addressof(dst), C.byref(sz), order)
    return dst.raw[:sz.value] if r else None


def rans_compress_4x16(data, order):
    """out == NULL form: the library malloc()s, we free()."""
    L = lib()
    data = bytes(data)
    src = C.create_string_buffer(data, max(len(data), 1))
    sz = C.c_uint(0)
    r = L.rans_compress_4x16(C.addressof(src), len(data), C.byref(sz), order)
    if not r:
        return None
    out = C.string_at(r, sz.value)
    _libc.free(r)
    return out


def rans_uncompress_to_4x16(comp, ulen):
    L = lib()
    comp = bytes(comp)
    src = C.create_string_buffer(comp, max(len(comp), 1))
    dst = C.create_string_buffer(max(ulen, 1))
    sz = C.c_uint(ulen)
    r = L.rans_uncompress_to_4x16(C.addressof(src), len(comp), C.addressof(dst), C.byref(sz))
    return dst.raw[:sz.value] if r else None


def rans_uncompress_4x16(comp):
    L = lib()
    comp = bytes(comp)
    src = C.create_string_buffer(comp, max(len(comp), 1))
    sz = C.c_uint(0)
    r = L.rans_uncompress_4x16(C.addressof(src), len(comp), C.byref(sz))
    if not r:
        return None
    out = C.string_at(r, sz.value)
    _libc.free(r)
    return out


# ---------------------------------------------------------------- batched host API
class PinnedBuffer:
    """Page-locked host memory from the library, viewed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().b200rans_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise B200RansError("pinned allocation of %d bytes failed" % nbytes)
        self.array = np.ctypeslib.as_array((C.c_ubyte * max(self.nbytes, 1)).from_address(self.ptr))[:self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            lib().b200rans_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _addr(a):
    return a.ctypes.data


def compress_batch(buf, offsets, sizes, orders, out=None, ngpu=1, block_of=None, multi=False):
    """Compress n slices of one host array.

    buf: uint8 numpy array (pinned for full PCIe speed); offsets/sizes/orders: per stream.
    Returns (out_array, out_off, out_size); stream k is out_array[out_off[k]:out_off[k]+out_size[k]].
    """
    L = lib()
    n = len(sizes)
    base = _addr(buf)
    ptrs = (np.asarray(offsets, np.uint64) + np.uint64(base)).astype(np.uint64)
    sizes = np.ascontiguousarray(sizes, np.uint32)
    orders = np.ascontiguousarray(orders, np.int32)
    if out is None:
        cap = int(L.b200rans_compress_batch_dev_bound(n, _addr(sizes), _addr(orders))) + 1024 * max(ngpu, 1)
        out = np.empty(cap, np.uint8)
    out_off = np.zeros(n, np.uint64)
    out_size = np.zeros(n, np.uint32)
    if ngpu == 1 and not multi:
        rc = L.b200rans_compress_batch(n, _addr(ptrs), _addr(sizes), _addr(orders), _addr(out), out.size,
                                       _addr(out_off), _addr(out_size))
    else:
        bo = None if block_of is None else np.ascontiguousarray(block_of, np.int32)
        rc = L.b200rans_compress_batch_multi(ngpu, n, _addr(ptrs), _addr(sizes), _addr(orders),
                                             _addr(bo) if bo is not None else None, _addr(out), out.size,
                                             _addr(out_off), _addr(out_size))
    _check(rc, "b200rans_compress_batch")
    return out, out_off, out_size


# rANS members of fqzcomp5's method enum (fqzcomp5.c:120-131) -> `order` argument
# (fqzcomp5.c:2005-2022); RANSXN1 needs the block's fixed read length.
RANS_METHOD_ORDERS = {"RANS0": 0, "RANS1": 1, "RANS64": 64, "RANS65": 65,
                      "RANS128": 128, "RANS129": 129, "RANS192": 192, "RANS193": 193}


def ransxn1_order(fixed_len):
    """fqzcomp5.c:2019: (fq->fixed_len << 8) + 9."""
    return (int(fixed_len) << 8) + 9


def compress_methods_batch(buf, offsets, sizes, methods, out=None):
    """Method trial (fqzcomp5.c:1979-2119, rANS members): every slice of `buf` is encoded
    under each `order` value in `methods`; the first smallest stream of each is returned.

    Returns (out_array, out_off, out_size, best, csize) with best[k] the index into
    `methods` and csize[k, j] the size under method j (0 = that call failed)."""
    L = lib()
    n = len(sizes)
    M = len(methods)
    ptrs = (np.asarray(offsets, np.uint64) + np.uint64(_addr(buf))).astype(np.uint64)
    sizes = np.ascontiguousarray(sizes, np.uint32)
    meth = np.ascontiguousarray(methods, np.int32)
    if out is None:
        cap = 256
        for k in range(n):
            cap += max(rans_compress_bound_4x16(int(sizes[k]), int(m)) for m in meth) + 32
        out = np.empty(cap, np.uint8)
    out_off = np.zeros(n, np.uint64)
    out_size = np.zeros(n, np.uint32)
    best = np.full(n, -1, np.int32)
    csize = np.zeros((n, M), np.uint32)
    rc = L.b200rans_compress_methods_batch(n, _addr(ptrs), _addr(sizes), M, _addr(meth), _addr(out), out.size,
                                           _addr(out_off), _addr(out_size), _addr(best), _addr(csize))
    _check(rc, "b200rans_compress_methods_batch")
    return out, out_off, out_size, best, csize


def compress_trials(buf, offsets, sizes, method_lists, out=None):
    """Ragged method trial: slice k of `buf` is tried under method_lists[k] (a list of `order`
    values).  Returns (out_array, out_off, out_size, best, csize_lists)."""
    L = lib()
    n = len(sizes)
    ptrs = (np.asarray(offsets, np.uint64) + np.uint64(_addr(buf))).astype(np.uint64)
    sizes = np.ascontiguousarray(sizes, np.uint32)
    first = np.zeros(n + 1, np.uint32)
    first[1:] = np.cumsum([len(m) for m in method_lists])
    flat = np.ascontiguousarray([o for m in method_lists for o in m], np.int32)
    if out is None:
        cap = 256
        for k in range(n):
            cap += max(rans_compress_bound_4x16(int(sizes[k]), int(m)) for m in method_lists[k]) + 32
        out = np.empty(cap, np.uint8)
    out_off = np.zeros(n, np.uint64)
    out_size = np.zeros(n, np.uint32)
    best = np.full(n, -1, np.int32)
    csize = np.zeros(len(flat), np.uint32)
    rc = L.b200rans_compress_trials(n, _addr(ptrs), _addr(sizes), _addr(first), _addr(flat), _addr(out), out.size,
                                    _addr(out_off), _addr(out_size), _addr(best), _addr(csize))
    _check(rc, "b200rans_compress_trials")
    return out, out_off, out_size, best, [csize[first[k]:first[k + 1]].tolist() for k in range(n)]


# tok3's per-token-type method tables (tokenise_name3.c:1283-1357, rANS build: bit 0x04 cleared at
# :1374-1375); levels 1-9 map to rows 0-4 by (level-1)//2.  Token types in enum order
# (tokenise_name3.c:96-112): TYPE ALPHA CHAR DIGITS0 DZLEN DUP DIFF DIGITS DDELTA DDELTA0 MATCH NOP END.
TOK3_TYPES = ["TYPE", "ALPHA", "CHAR", "DIGITS0", "DZLEN", "DUP", "DIFF", "DIGITS", "DDELTA", "DDELTA0",
              "MATCH", "NOP", "END"]


# R[level][type] of tokenise_name3.c:1283-1357, one row per level group (-1, -3, -5, -7, -9)
TOK3_METHODS = [
    [[128], [129], [0], [8], [0], [8], [8], [8], [0], [128], [0], [0], [0]],
    [[192, 0], [129, 1], [0], [136, 0], [0], [200], [136], [200], [0], [128], [0], [0], [0]],
    [[192, 0], [1, 128, 0, 129], [0], [200, 0], [0], [200], [192, 200], [132, 201], [0], [128], [0], [0], [0]],
    [[193, 0, 1], [128, 1, 128, 0, 129], [1, 0], [200, 0], [0], [201], [192, 200], [132, 201], [0], [128], [0],
     [0], [0]],
    [[192, 0, 1, 65, 193, 132], [132, 1, 0, 129], [1, 0, 192], [201, 0, 192, 64], [0, 128, 1], [201],
     [192, 201, 65], [132, 201, 1, 192, 129, 193], [1, 0, 192], [192, 1, 0], [0], [0], [0]],
]


def tok3_level_row(level):
    """tokenise_name3.c:1275-1278: levels 1-9 -> rows 0-4."""
    return min(4, max(0, (level - 1) // 2))


def tok3_method_list(table_row, in_len):
    """The candidate list tok3's compress() walks for one token stream: the row of its R[level][type]
    table with X32 cleared (:1374-1375) and STRIPE entries dropped when in_len % 4 != 0 (:1377-1378)."""
    out = []
    for m in table_row:
        m &= ~4
        if in_len % 4 != 0 and (m & 8):
            continue
        out.append(m)
    return out


def compress_methods(data, methods):
    """Single buffer, malloc()ed winner: (bytes or None, best, csize)."""
    L = lib()
    data = bytes(data)
    src = C.create_string_buffer(data, max(len(data), 1))
    meth = np.ascontiguousarray(methods, np.int32)
    csize = np.zeros(len(meth), np.uint32)
    sz = C.c_uint(0)
    best = C.c_int(-1)
    r = L.b200rans_compress_methods(C.addressof(src), len(data), len(meth), _addr(meth), C.byref(sz),
                                    C.byref(best), _addr(csize))
    if not r:
        return None, best.value, csize
    out = C.string_at(r, sz.value)
    _libc.free(r)
    return out, best.value, csize


def uncompress_batch(comp, comp_off, comp_size, out, out_off, out_size, ngpu=1, block_of=None, multi=False):
    """Decompress n streams held in one host array into slices of `out`.

    out_size[k] is the capacity (exact length for NOSZ streams).  Returns (sizes, status)."""
    L = lib()
    n = len(comp_size)
    iptr = (np.asarray(comp_off, np.uint64) + np.uint64(_addr(comp))).astype(np.uint64)
    optr = (np.asarray(out_off, np.uint64) + np.uint64(_addr(out))).astype(np.uint64)
    isz = np.ascontiguousarray(comp_size, np.uint32)
    osz = np.array(out_size, np.uint32)
    status = np.zeros(n, np.int32)
    if ngpu == 1 and not multi:
        rc = L.b200rans_uncompress_batch(n, _addr(iptr), _addr(isz), _addr(optr), _addr(osz), _addr(status))
    else:
        bo = None if block_of is None else np.ascontiguousarray(block_of, np.int32)
        rc = L.b200rans_uncompress_batch_multi(ngpu, n, _addr(iptr), _addr(isz),
                                               _addr(bo) if bo is not None else None, _addr(optr), _addr(osz),
                                               _addr(status))
    _check(rc, "b200rans_uncompress_batch")
    return osz, status


def uncompressed_size(comp):
    comp = bytes(comp)
    return int(lib().b200rans_uncompressed_size(comp, len(comp)))


# ---------------------------------------------------------------- device-resident API
def compress_batch_dev(stream, d_in_ptr, in_off, in_size, orders, d_out_ptr, out_cap, d_out_off_ptr,
                       d_out_size_ptr):
    """Pointers are raw device addresses (e.g. torch.Tensor.data_ptr()); descriptor arrays are numpy."""
    in_off = np.ascontiguousarray(in_off, np.uint64)
    in_size = np.ascontiguousarray(in_size, np.uint32)
    orders = np.ascontiguousarray(orders, np.int32)
    rc = lib().b200rans_compress_batch_dev(stream, len(in_size), d_in_ptr, _addr(in_off), _addr(in_size),
                                           _addr(orders), d_out_ptr, out_cap, d_out_off_ptr, d_out_size_ptr)
    _check(rc, "b200rans_compress_batch_dev")


def uncompress_batch_dev(stream, d_in_ptr, in_off, in_size, d_out_ptr, out_off, out_size, d_out_size_ptr,
                         d_status_ptr, flags=None):
    in_off = np.ascontiguousarray(in_off, np.uint64)
    in_size = np.ascontiguousarray(in_size, np.uint32)
    out_off = np.ascontiguousarray(out_off, np.uint64)
    out_size = np.ascontiguousarray(out_size, np.uint32)
    if flags is not None:
        flags = np.ascontiguousarray(flags, np.uint8)
    rc = lib().b200rans_uncompress_batch_dev(stream, len(in_size), d_in_ptr, _addr(in_off), _addr(in_size),
                                             _addr(flags) if flags is not None else None,
                                             d_out_ptr, _addr(out_off), _addr(out_size), d_out_size_ptr,
                                             d_status_ptr)
    _check(rc, "b200rans_uncompress_batch_dev")


OUT_IN_SLOT = 1


def compress_slots_bound(in_size, orders):
    """d_out bytes b200rans_compress_batch_dev2 needs with OUT_IN_SLOT (one bound-sized slot per call)."""
    in_size = np.ascontiguousarray(in_size, np.uint32)
    orders = np.ascontiguousarray(orders, np.int32)
    return int(lib().b200rans_compress_slots_bound(len(in_size), _addr(in_size), _addr(orders)))


def compress_batch_dev2(stream, d_in_ptr, in_off, in_size, orders, d_out_ptr, out_cap, d_out_off_ptr,
                        d_out_size_ptr, flags=0):
    """As compress_batch_dev; flags=OUT_IN_SLOT leaves every stream in its own slot of d_out."""
    in_off = np.ascontiguousarray(in_off, np.uint64)
    in_size = np.ascontiguousarray(in_size, np.uint32)
    orders = np.ascontiguousarray(orders, np.int32)
    rc = lib().b200rans_compress_batch_dev2(stream, len(in_size), d_in_ptr, _addr(in_off), _addr(in_size),
                                            _addr(orders), d_out_ptr, out_cap, d_out_off_ptr, d_out_size_ptr, flags)
    _check(rc, "b200rans_compress_batch_dev2")


def compress_trials_dev(stream, d_in_ptr, in_off, in_size, method_lists, d_out_ptr, out_cap, pack_align,
                        d_out_off_ptr, d_out_size_ptr, d_best_ptr=None, d_csize_ptr=None):
    """Ragged method trial with inputs and outputs in HBM (b200rans_compress_trials_dev)."""
    in_off = np.ascontiguousarray(in_off, np.uint64)
    in_size = np.ascontiguousarray(in_size, np.uint32)
    n = len(in_size)
    first = np.zeros(n + 1, np.uint32)
    first[1:] = np.cumsum([len(m) for m in method_lists])
    flat = np.ascontiguousarray([o for m in method_lists for o in m], np.int32)
    rc = lib().b200rans_compress_trials_dev(stream, n, d_in_ptr, _addr(in_off), _addr(in_size), _addr(first),
                                            _addr(flat), d_out_ptr, out_cap, pack_align, d_out_off_ptr,
                                            d_out_size_ptr, d_best_ptr, d_csize_ptr)
    _check(rc, "b200rans_compress_trials_dev")
    return first


def tok3_methods(level, token_type, in_len):
    """b200rans_tok3_methods: the candidate list tok3's compress() walks (C-side tables)."""
    out = np.zeros(8, np.int32)
    n = lib().b200rans_tok3_methods(level, token_type, in_len, _addr(out))
    if n < 0:
        raise B200RansError("b200rans_tok3_methods: bad arguments")
    return out[:n].tolist()


def compress_bound_batch(in_size, orders):
    in_size = np.ascontiguousarray(in_size, np.uint32)
    orders = np.ascontiguousarray(orders, np.int32)
    return int(lib().b200rans_compress_batch_dev_bound(len(in_size), _addr(in_size), _addr(orders)))


def launch_count():
    return int(lib().b200rans_launch_count())


def set_profiling(on):
    _check(lib().b200rans_set_profiling(1 if on else 0), "b200rans_set_profiling")


def dec_staged_stats(reset=False):
    """Counters of the staged decode's head stage: [taken, failed after taken, handed back by reason ...]."""
    out = (C.c_ulonglong * 16)()
    _check(lib().b200rans_dec_staged_stats(out, 1 if reset else 0), "b200rans_dec_staged_stats")
    return [int(x) for x in out]


def last_kernel_ms(which):
    """which: 0 = encode coder kernel, 1 = decode coder kernel of the last batch call."""
    return float(lib().b200rans_last_kernel_ms(which))


# ---------------------------------------------------------------- FASTQ split / join
class FqInfo(C.Structure):
    """b200fq_info (include/b200rans.h)."""
    _fields_ = [("status", C.c_int32), ("num_records", C.c_uint32), ("name_len", C.c_uint32),
                ("seq_len", C.c_uint32), ("qual_len", C.c_uint32), ("fixed_len", C.c_int32),
                ("consumed", C.c_uint32), ("text_len", C.c_uint32), ("more", C.c_uint32)]


def load_seqs_kseq(text, blk_size, max_records=None):
    """The live loader's rules (load_seqs_kseq, fqzcomp5.c:423-623) on the GPU for strict 4-line FASTQ: one
    block of at most blk_size accounted bytes from the front of `text`.  Returns load_seqs' dict plus `more`
    (1: the block-size rule ended the block; 0: the text ran out), or None where the reference fails."""
    return load_seqs(text, max_records, mode=1, blk_size=blk_size)


def load_seqs(text, max_records=None, mode=0, blk_size=0):
    """The reference's load_seqs (fqzcomp5.c:279-410) on the GPU: a block of FASTQ text ->
    dict(num_records, name, seq, qual, len, flag, fixed_len, consumed), or None where the
    reference returns NULL.  `text`: bytes or a uint8 numpy array (pinned for full PCIe rate)."""
    L = lib()
    t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text
    n = int(t.size)
    mr = int(max_records) if max_records is not None else n // 24 + 64
    while True:
        name = np.empty(n + 16, np.uint8); seq = np.empty(n + 16, np.uint8); qual = np.empty(n + 16, np.uint8)
        ln = np.empty(mr + 1, np.uint32); fl = np.empty(mr + 1, np.uint32)
        info = FqInfo()
        rc = L.b200fq_split_mode(mode, blk_size, _addr(t) if n else None, n, _addr(name), n + 16, _addr(seq),
                                 _addr(qual), n + 16, _addr(ln), _addr(fl), mr, C.addressof(info))
        _check(rc, "b200fq_split")
        if info.status == 2 and max_records is None and mr < n // 4 + 64:
            mr = n // 4 + 64            # more records than guessed: the smallest record is 6 bytes... retry
            continue
        break
    if info.status:
        if info.status == 2:
            raise B200RansError("b200fq_split: max_records too small")
        return None
    R = info.num_records
    return dict(num_records=R, name=name[:info.name_len].tobytes(), seq=seq[:info.seq_len].tobytes(),
                qual=qual[:info.qual_len].tobytes(), len=ln[:R].tolist(), flag=fl[:R].tolist(),
                fixed_len=info.fixed_len, consumed=info.consumed, **({"more": info.more} if mode else {}))


def output_fastq(name, seq, qual, lens, plus_name=0):
    """The reference's output_fastq (fqzcomp5.c:3440-3480) with the decoder's +33 on the
    qualities (fqzcomp5.c:2532-2533), on the GPU.  Returns the FASTQ text as bytes."""
    L = lib()
    nb = np.frombuffer(bytes(name), np.uint8); sb = np.frombuffer(bytes(seq), np.uint8)
    qb = np.frombuffer(bytes(qual), np.uint8)
    ln = np.ascontiguousarray(lens, np.uint32)
    R = int(ln.size)
    cap = 2 * nb.size + 2 * sb.size + 6 * R + 64
    out = np.empty(cap, np.uint8)
    info = FqInfo()
    rc = L.b200fq_join(_addr(nb) if nb.size else None, nb.size, _addr(sb) if sb.size else None,
                       _addr(qb) if qb.size else None, sb.size, _addr(ln) if R else None, R, int(plus_name),
                       _addr(out), cap, C.addressof(info))
    _check(rc, "b200fq_join")
    if info.status:
        return None
    return out[:info.text_len].tobytes()


# ---------------------------------------------------------------- CRC-32 and block framing
class FqzPiece(C.Structure):
    """b200fqz_piece (include/b200rans.h)."""
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_uint32), ("on_device", C.c_int)]


def crc32(data, crc_in=0):
    """zlib's crc32(crc_in, data) computed on the GPU (host buffer in)."""
    b = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data
    out = C.c_uint(0)
    _check(lib().b200fqz_crc32(crc_in, _addr(b) if b.size else None, b.size, C.byref(out)), "b200fqz_crc32")
    return out.value


def assemble_block_dev(stream, num_records, pieces, d_block_ptr, block_cap):
    """pieces: list of (address, length, on_device).  Returns the block length."""
    arr = (FqzPiece * max(len(pieces), 1))()
    for i, (ptr, ln, dev) in enumerate(pieces):
        arr[i].ptr, arr[i].len, arr[i].on_device = ptr, ln, int(dev)
    n = C.c_uint(0)
    _check(lib().b200fqz_assemble_block_dev(stream, num_records, len(pieces), C.addressof(arr), d_block_ptr,
                                            block_cap, C.byref(n)), "b200fqz_assemble_block_dev")
    return n.value


# ---------------------------------------------------------------- one call per fqzcomp5 block (part 5)
RANSXN1 = -1            # B200FQZ_RANSXN1: (fixed_len << 8) + 9, skipped for variable-length reads
STRAT_SLICED = 0xB2
MAX_METHODS = 16
NAME_CODER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32,
                         C.POINTER(C.c_void_p), C.POINTER(C.c_uint32))


class BlockOpts(C.Structure):
    """b200fqz_block_opts."""
    _fields_ = [("slice_bytes", C.c_uint32), ("n_name_methods", C.c_int), ("n_seq_methods", C.c_int),
                ("n_qual_methods", C.c_int), ("name_methods", C.c_int * MAX_METHODS),
                ("seq_methods", C.c_int * MAX_METHODS), ("qual_methods", C.c_int * MAX_METHODS),
                ("name_coder", NAME_CODER), ("name_user", C.c_void_p), ("kseq_blk_size", C.c_uint32)]


class BlockReport(C.Structure):
    """b200fqz_block_report."""
    _fields_ = [("status", C.c_int32), ("num_records", C.c_uint32), ("consumed", C.c_uint32),
                ("fixed_len", C.c_int32), ("ulen", C.c_uint32 * 3), ("clen", C.c_uint32 * 3),
                ("nslices", C.c_uint32 * 3), ("csize", (C.c_uint64 * MAX_METHODS) * 3),
                ("wins", (C.c_uint32 * MAX_METHODS) * 3), ("block_len", C.c_uint32), ("crc", C.c_uint32),
                ("ms", C.c_float * 4)]


class Learner(C.Structure):
    """b200fqz_learner: metrics_method / metrics_update (fqzcomp5.c:1899-1958) per section."""
    _fields_ = [("review", C.c_int32 * 3), ("trial", C.c_int32 * 3), ("used", C.c_int32 * 3),
                ("usize", (C.c_uint64 * MAX_METHODS) * 3), ("csize", (C.c_uint64 * MAX_METHODS) * 3),
                ("on_trial", C.c_int32 * 3)]

    def __init__(self):
        super().__init__()
        lib().b200fqz_learner_init(C.addressof(self))

    def methods(self, all_opts):
        """The options this block is coded with (every method while a trial is on, else the best one)."""
        out = BlockOpts()
        lib().b200fqz_learner_methods(C.addressof(self), C.addressof(all_opts), C.addressof(out))
        return out

    def update(self, used_opts, rep):
        lib().b200fqz_learner_update(C.addressof(self), C.addressof(used_opts), C.addressof(rep))


# fqzcomp5 -3 (fqzcomp5.c:4893-4900), the rANS members of its seq / qual method sets; names: the codec half
# of TLZP3 (order 5).  x32=True ORs RANS_ORDER_X32 into every method.
def block_opts(slice_bytes=262144, seq=(0, 1, 129, 193), qual=(0, 1, 129, 193, RANSXN1), names=(5,), x32=False,
               name_coder=None):
    o = BlockOpts()
    o.slice_bytes = slice_bytes
    fix = lambda m: m if m == RANSXN1 or not x32 else (m | RANS_ORDER_X32)
    for lst, arr, cnt in ((names, o.name_methods, "n_name_methods"), (seq, o.seq_methods, "n_seq_methods"),
                          (qual, o.qual_methods, "n_qual_methods")):
        for i, m in enumerate(lst):
            arr[i] = fix(m)
        setattr(o, cnt, len(lst))
    if name_coder is not None:
        o.name_coder = name_coder
    return o


def resolve_methods(lst, fixed_len, x32=False):
    """The `order` values a method list stands for in a block whose reads are fixed_len long."""
    out = []
    for m in lst:
        if m == RANSXN1:
            if fixed_len <= 0:
                continue
            out.append((fixed_len << 8) + 9)
        else:
            out.append(m | (RANS_ORDER_X32 if x32 else 0))
    return out


def encode_block(text, opts, out=None):
    """b200fqz_encode_block: FASTQ text (bytes or uint8 array) -> (block bytes view, BlockReport)."""
    t = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text
    n = int(t.size)
    cap = int(lib().b200fqz_block_bound(n))
    if out is None:
        out = np.empty(cap, np.uint8)
    rep = BlockReport()
    rc = lib().b200fqz_encode_block(_addr(t) if n else None, n, C.addressof(opts), _addr(out), out.size,
                                    C.addressof(rep))
    _check(rc, "b200fqz_encode_block")
    return out[:rep.block_len], rep


def decode_block(block, text_cap, plus_name=0, out=None):
    """b200fqz_decode_block -> (text bytes view or None, BlockReport)."""
    b = np.frombuffer(block, np.uint8) if isinstance(block, (bytes, bytearray)) else block
    if out is None:
        out = np.empty(text_cap, np.uint8)
    rep = BlockReport()
    rc = lib().b200fqz_decode_block(_addr(b), int(b.size), plus_name, _addr(out), out.size, C.addressof(rep))
    _check(rc, "b200fqz_decode_block")
    return (out[:rep.block_len] if rep.status == 0 else None), rep


def encode_blocks_multi(ngpu, texts, opts, outs):
    """texts / outs: lists of uint8 arrays (pinned for full PCIe rate).  Returns the list of BlockReport."""
    nb = len(texts)
    tp = np.array([_addr(t) for t in texts], np.uint64)
    tn = np.array([t.size for t in texts], np.uint32)
    bp = np.array([_addr(b) for b in outs], np.uint64)
    bc = np.array([b.size for b in outs], np.uint64)
    reps = (BlockReport * max(nb, 1))()
    rc = lib().b200fqz_encode_blocks_multi(ngpu, nb, _addr(tp), _addr(tn), C.addressof(opts), _addr(bp), _addr(bc),
                                           C.addressof(reps))
    _check(rc, "b200fqz_encode_blocks_multi")
    return [reps[i] for i in range(nb)]


def decode_blocks_multi(ngpu, blocks, lens, outs, plus_name=0):
    nb = len(blocks)
    bp = np.array([_addr(b) for b in blocks], np.uint64)
    bl = np.array(lens, np.uint32)
    tp = np.array([_addr(t) for t in outs], np.uint64)
    tc = np.array([t.size for t in outs], np.uint64)
    reps = (BlockReport * max(nb, 1))()
    rc = lib().b200fqz_decode_blocks_multi(ngpu, nb, _addr(bp), _addr(bl), plus_name, _addr(tp), _addr(tc),
                                           C.addressof(reps))
    _check(rc, "b200fqz_decode_blocks_multi")
    return [reps[i] for i in range(nb)]
