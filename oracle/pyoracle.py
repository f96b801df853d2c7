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


# ---------------------------------------------------------------- FASTQ split / join checkers
_FQ_LIBS = {
    "oracle": os.path.join(HERE, "liboracle_fastq.so"),
    "ref": os.path.join(HERE, "_ref", "libref_fqz.so"),
}


def fastq_available(kind):
    return os.path.exists(_FQ_LIBS[kind])


class FqInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("num_records", C.c_uint32), ("name_len", C.c_uint32),
                ("seq_len", C.c_uint32), ("qual_len", C.c_uint32), ("fixed_len", C.c_int32),
                ("consumed", C.c_uint32), ("text_len", C.c_uint32), ("more", C.c_uint32)]


class _RefFastq(C.Structure):          # `fastq`, fqzcomp5.c:235-249
    _fields_ = [("num_records", C.c_int), ("name_buf", C.c_void_p), ("seq_buf", C.c_void_p),
                ("qual_buf", C.c_void_p), ("name", C.POINTER(C.c_int)), ("seq", C.POINTER(C.c_int)),
                ("qual", C.POINTER(C.c_int)), ("len", C.POINTER(C.c_uint)), ("flag", C.POINTER(C.c_uint)),
                ("name_len", C.c_int), ("seq_len", C.c_int), ("qual_len", C.c_int),
                ("name_sz", C.c_int), ("seq_sz", C.c_int), ("qual_sz", C.c_int),
                ("fixed_len", C.c_int), ("is_fasta", C.c_int)]


class FastqChecker:
    """split(text) -> dict or None (the reference returns NULL); join(...) -> bytes.

    kind "oracle": oracle/fastq_oracle.c.  kind "ref": the reference's own load_seqs /
    output_fastq (fqzcomp5.c:279-410, :3440-3480) from oracle/_ref/libref_fqz.so; the
    +33 the decoder applies before output_fastq (fqzcomp5.c:2532-2533) is applied here."""

    def __init__(self, kind="oracle"):
        path = _FQ_LIBS[kind]
        if not os.path.exists(path) and (kind == "oracle" or os.path.isdir(REFERENCE_ROOT)):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.kind = kind
        self.lib = L = C.CDLL(path)
        if kind == "oracle":
            L.fqo_split.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 5 + [C.POINTER(FqInfo)]
            L.fqo_join.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_int, C.c_void_p]
            L.fqo_join.restype = C.c_uint32
            L.fqo_split_kseq.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5 + [C.POINTER(FqInfo)]
        else:
            L.load_seqs.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
            L.load_seqs.restype = C.POINTER(_RefFastq)
            L.fastq_free.argtypes = [C.POINTER(_RefFastq)]
            L.output_fastq.argtypes = [C.c_void_p, C.POINTER(_RefFastq), C.c_int]
            L.load_seqs_kseq.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
            L.load_seqs_kseq.restype = C.POINTER(_RefFastq)
            self.z = C.CDLL("libz.so.1")
            self.z.gzopen.argtypes = [C.c_char_p, C.c_char_p]
            self.z.gzopen.restype = C.c_void_p
            self.z.gzclose.argtypes = [C.c_void_p]
            _libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
            _libc.fopen.restype = C.c_void_p
            _libc.fclose.argtypes = [C.c_void_p]

    def split(self, text):
        import numpy as np
        text = bytes(text)
        n = len(text)
        src = C.create_string_buffer(text, max(n, 1))
        if self.kind == "oracle":
            name = np.zeros(n + 16, np.uint8); seq = np.zeros(n + 16, np.uint8); qual = np.zeros(n + 16, np.uint8)
            ln = np.zeros(n // 4 + 16, np.uint32); fl = np.zeros(n // 4 + 16, np.uint32)
            info = FqInfo()
            self.lib.fqo_split(C.addressof(src), n, name.ctypes.data, seq.ctypes.data, qual.ctypes.data,
                               ln.ctypes.data, fl.ctypes.data, C.byref(info))
            if info.status:
                return None
            R = info.num_records
            return dict(num_records=R, name=name[:info.name_len].tobytes(), seq=seq[:info.seq_len].tobytes(),
                        qual=qual[:info.qual_len].tobytes(), len=ln[:R].tolist(), flag=fl[:R].tolist(),
                        fixed_len=info.fixed_len, consumed=info.consumed)
        last = C.c_int(0)
        devnull = os.open(os.devnull, os.O_WRONLY)       # load_seqs reports errors on stderr
        saved = os.dup(2)
        os.dup2(devnull, 2)
        try:
            fq = self.lib.load_seqs(C.addressof(src), n, C.byref(last))
        finally:
            os.dup2(saved, 2); os.close(saved); os.close(devnull)
        if not fq:
            return None
        f = fq.contents
        R = f.num_records
        out = dict(num_records=R, name=C.string_at(f.name_buf, f.name_len), seq=C.string_at(f.seq_buf, f.seq_len),
                   qual=C.string_at(f.qual_buf, f.qual_len), len=[f.len[i] for i in range(R)],
                   flag=[f.flag[i] for i in range(R)], fixed_len=f.fixed_len, consumed=last.value)
        self.lib.fastq_free(fq)
        return out

    def split_kseq(self, text, blk_size):
        """The blocks the live loader (load_seqs_kseq, fqzcomp5.c:423-623) cuts `text` into with this
        blk_size: a list of load_seqs-style dicts (`consumed` only from the oracle), or None where it fails."""
        import numpy as np
        import tempfile
        text = bytes(text)
        n = len(text)
        blocks = []
        if self.kind == "oracle":
            src = C.create_string_buffer(text, max(n, 1))
            pos = 0
            while True:
                m = n - pos
                name = np.zeros(m + 16, np.uint8); seq = np.zeros(m + 16, np.uint8); qual = np.zeros(m + 16, np.uint8)
                ln = np.zeros(m // 4 + 16, np.uint32); fl = np.zeros(m // 4 + 16, np.uint32)
                info = FqInfo()
                self.lib.fqo_split_kseq(C.addressof(src) + pos, m, blk_size, name.ctypes.data, seq.ctypes.data,
                                        qual.ctypes.data, ln.ctypes.data, fl.ctypes.data, C.byref(info))
                if info.status:
                    return None
                R = info.num_records
                if R == 0:
                    break
                blocks.append(dict(num_records=R, name=name[:info.name_len].tobytes(),
                                   seq=seq[:info.seq_len].tobytes(), qual=qual[:info.qual_len].tobytes(),
                                   len=ln[:R].tolist(), flag=fl[:R].tolist(), fixed_len=info.fixed_len,
                                   consumed=info.consumed, more=info.more))
                pos += info.consumed
                if not info.more:
                    break
            return blocks
        with tempfile.NamedTemporaryFile(delete=False) as t:
            t.write(text)
            path = t.name
        fp = self.z.gzopen(path.encode(), b"rb")
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)
        try:
            while True:
                eof = C.c_int(0)
                fq = self.lib.load_seqs_kseq(fp, blk_size, C.byref(eof))
                if not fq:
                    blocks = None
                    break
                f = fq.contents
                R = f.num_records
                if R:
                    blocks.append(dict(num_records=R, name=C.string_at(f.name_buf, f.name_len),
                                       seq=C.string_at(f.seq_buf, f.seq_len), qual=C.string_at(f.qual_buf, f.qual_len),
                                       len=[f.len[i] for i in range(R)], flag=[f.flag[i] for i in range(R)],
                                       fixed_len=f.fixed_len))
                self.lib.fastq_free(fq)
                if eof.value or not R:
                    break
        finally:
            os.dup2(saved, 2); os.close(saved); os.close(devnull)
            self.z.gzclose(fp)
            os.unlink(path)
        return blocks

    def join(self, name, seq, qual, lens, plus_name=0):
        import numpy as np
        import tempfile
        R = len(lens)
        ln = np.ascontiguousarray(lens, np.uint32)
        nb = C.create_string_buffer(bytes(name), max(len(name), 1))
        sb = C.create_string_buffer(bytes(seq), max(len(seq), 1))
        if self.kind == "oracle":
            qb = C.create_string_buffer(bytes(qual), max(len(qual), 1))
            out = np.zeros(2 * len(name) + 2 * len(seq) + 6 * R + 16, np.uint8)
            k = self.lib.fqo_join(C.addressof(nb), C.addressof(sb), C.addressof(qb), ln.ctypes.data, R, plus_name,
                                  out.ctypes.data)
            return out[:k].tobytes()
        q33 = bytes((b + 33) & 0xff for b in bytes(qual))           # fqzcomp5.c:2532-2533
        qb = C.create_string_buffer(q33, max(len(q33), 1))
        fq = _RefFastq()
        fq.num_records = R
        fq.name_buf, fq.seq_buf, fq.qual_buf = C.addressof(nb), C.addressof(sb), C.addressof(qb)
        fq.len = ln.ctypes.data_as(C.POINTER(C.c_uint))
        fq.name_len, fq.seq_len, fq.qual_len = len(name), len(seq), len(qual)
        with tempfile.NamedTemporaryFile(delete=False) as t:
            path = t.name
        fp = _libc.fopen(path.encode(), b"wb")
        self.lib.output_fastq(fp, C.byref(fq), plus_name)
        _libc.fclose(fp)
        data = open(path, "rb").read()
        os.unlink(path)
        return data


# ---------------------------------------------------------------- CRC-32 / block framing checker
class Crc32Oracle:
    """oracle/crc32_oracle.c: crc32(data, crc) and frame_block(num_records, pieces)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle_crc32.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.orc_crc32.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64]
        L.orc_crc32.restype = C.c_uint32
        L.orc_frame_block.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_frame_block.restype = C.c_uint32

    def crc32(self, data, crc=0):
        data = bytes(data)
        b = C.create_string_buffer(data, max(len(data), 1))
        return int(self.lib.orc_crc32(crc, C.addressof(b), len(data)))

    def frame_block(self, num_records, pieces):
        bufs = [C.create_string_buffer(bytes(p), max(len(p), 1)) for p in pieces]
        ptrs = (C.c_void_p * max(len(bufs), 1))(*[C.addressof(b) for b in bufs])
        lens = (C.c_uint32 * max(len(bufs), 1))(*[len(p) for p in pieces])
        out = C.create_string_buffer(12 + sum(len(p) for p in pieces) + 16)
        n = self.lib.orc_frame_block(num_records, len(pieces), C.addressof(ptrs), C.addressof(lens), C.addressof(out))
        return out.raw[:n]
