"""One call per fqzcomp5 block (include/b200rans.h part 5): b200fqz_encode_block /
b200fqz_decode_block against the reference's own pieces.

encode_block (fqzcomp5.c:2147-2280) = load_seqs + compress_with_methods per section + framing + CRC.
Parity is checked piece by piece: the records and field buffers equal load_seqs' (checked through the
slices' decoded bytes), every slice's stream equals rans_compress_4x16(slice, methods[winner]) with
the winner of the serial CPU loop, the per-method size sums equal the CPU loop's, the CRC equals
zlib's over the same bytes, and decode(encode(text)) is the text load_seqs consumed.
"""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest

import corpus
from fqzcomp5_b200 import synth

pytestmark = pytest.mark.gpu


def fastq_text(nrec, rl=150, seed=1, variable=False, pair=False):
    rng = np.random.default_rng(seed)
    seq = synth.illumina_seq(nrec * rl, seed=seed + 3).reshape(nrec, rl)
    qual = (synth.illumina_qual(nrec * rl, seed=seed + 2) + 33).astype(np.uint8).reshape(nrec, rl)
    out = []
    for i in range(nrec):
        ln = int(rng.integers(30, rl + 1)) if variable else rl
        name = b"@SIM.%d %d/%d" % (seed * 1000000 + i // (2 if pair else 1), i, (i % 2 + 1) if pair else 1)
        out.append(name + b"\n" + seq[i, :ln].tobytes() + b"\n+\n" + qual[i, :ln].tobytes() + b"\n")
    return b"".join(out)


def _cpu_trial(checker, data, methods):
    best, best_out, sizes = -1, None, []
    for j, order in enumerate(methods):
        out = checker.compress_malloc(data, order)
        sizes.append(len(out) if out else 0)
        if out and (best_out is None or len(best_out) > len(out)):
            best, best_out = j, out
    return best, best_out, sizes


def parse_block(block):
    """The framing of include/b200rans.h part 5 -> dict(num_records, crc, sections=[(strat, ulen, S, [streams])], len_piece)."""
    b = bytes(block)
    size, nrec, crc = struct.unpack_from("<III", b, 0)
    assert size == len(b) - 4
    p = 12
    secs = []
    len_piece = None
    for s in range(3):
        if s == 0:
            ulen, strat, clen = struct.unpack_from("<IBI", b, p)
        else:
            strat, ulen, clen = struct.unpack_from("<BII", b, p)
        p += 9
        q, qend = p, p + clen
        if strat == 0xB2:
            nsl, S = struct.unpack_from("<II", b, q)
            q += 8
            cs = struct.unpack_from("<%dI" % nsl, b, q)
            q += 4 * nsl
            streams = []
            for c in cs:
                streams.append(b[q:q + c])
                q += c
            assert q == qend
        else:
            S, streams = max(ulen, 1), [b[q:qend]]
        secs.append((strat, ulen, S, streams))
        p = qend
        if s == 0:
            if b[p] > 0:
                len_piece = b[p:p + 1 + b[p]]
                p += 1 + b[p]
            else:
                (bl,) = struct.unpack_from("<I", b, p + 1)
                len_piece = b[p:p + 5 + bl]
                p += 5 + bl
    assert p == len(b)
    return dict(num_records=nrec, crc=crc, sections=secs, len_piece=len_piece)


def check_block(gpu_codec, checker, text, opts_kw, lists):
    """Encode, compare every piece with the CPU, decode, compare with the text."""
    from oracle.pyoracle import FastqChecker, fastq_available
    fqc = FastqChecker("ref" if fastq_available("ref") else "oracle")
    want = fqc.split(text)
    opts = gpu_codec.block_opts(**opts_kw)
    block, rep = gpu_codec.encode_block(text, opts)
    if want is None:
        assert rep.status == 1
        return None
    assert rep.status == 0 and rep.num_records == want["num_records"] and rep.consumed == want["consumed"]
    assert rep.fixed_len == want["fixed_len"]
    B = parse_block(block)
    assert B["num_records"] == want["num_records"]
    assert B["crc"] == zlib.crc32(bytes(block[12:])) == rep.crc          # fqzcomp5.c:2266-2274
    fields = [want["name"], want["seq"], want["qual"]]
    x32 = opts_kw.get("x32", False)
    for s, (strat, ulen, S, streams) in enumerate(B["sections"]):
        data = fields[s]
        assert ulen == len(data) == rep.ulen[s]
        methods = gpu_codec.resolve_methods(lists[s], want["fixed_len"], x32)
        sums = [0] * len(methods)
        wins = [0] * len(methods)
        assert len(streams) == max(1, -(-len(data) // S)) == rep.nslices[s]
        for i, got in enumerate(streams):
            piece = data[i * S:(i + 1) * S]
            wb, wout, wsizes = _cpu_trial(checker, piece, methods)
            assert got == wout, (s, i, hex(methods[wb]), len(got), len(wout))
            wins[wb] += 1
            sums = [a + b for a, b in zip(sums, wsizes)]
        assert list(rep.csize[s])[:len(methods)] == sums          # what metrics_update is fed (fqzcomp5.c:1950-1958)
        assert list(rep.wins[s])[:len(methods)] == wins
    if want["fixed_len"] > 0:                                      # fqzcomp5.c:2190-2197
        v = want["fixed_len"]
        enc = bytes([v]) if v < 128 else bytes([0x80 | (v >> 7), v & 0x7f])
        assert B["len_piece"] == bytes([len(enc)]) + enc
    back, drep = gpu_codec.decode_block(block, len(text) + 64)
    assert drep.status == 0 and back is not None
    assert back.tobytes() == text[:want["consumed"]]
    return block


LISTS = [(5,), (0, 1, 129, 193), (0, 1, 129, 193, -1)]             # names / seq / qual of fqzcomp5 -3 (:4893-4900)


@pytest.mark.parametrize("x32", [False, True])
def test_block_sliced_m3(gpu_codec, checker, x32):
    """fqzcomp5 -3 method sets over 256 KiB slices, stock 4-lane and with RANS_ORDER_X32."""
    text = fastq_text(6000, 150, seed=2)
    check_block(gpu_codec, checker, text, dict(slice_bytes=262144, x32=x32), LISTS)


def test_block_whole_sections(gpu_codec, checker):
    """slice_bytes = 0: one stream per section, the reference's own section layout (strat 0)."""
    text = fastq_text(1500, 150, seed=3)
    block = check_block(gpu_codec, checker, text, dict(slice_bytes=0), LISTS)
    B = parse_block(block)
    assert [s[0] for s in B["sections"]] == [0xB0, 0, 0]


def test_block_variable_lengths_and_cut_record(gpu_codec, checker):
    """Variable read lengths (RANSXN1 is skipped, lengths become varints) and a block that ends inside a record."""
    text = fastq_text(900, 120, seed=4, variable=True, pair=True)
    check_block(gpu_codec, checker, text[:-37], dict(slice_bytes=20000), LISTS)


def test_block_small_and_empty(gpu_codec, checker):
    for text in (b"", fastq_text(1, 50, seed=5), fastq_text(3, 7, seed=6)):
        check_block(gpu_codec, checker, text, dict(slice_bytes=4096), LISTS)


def test_block_malformed_and_corrupt(gpu_codec, checker):
    text = fastq_text(50, 100, seed=7)
    bad = text.replace(b"\n+\n", b"\n-\n", 1)
    opts = gpu_codec.block_opts(slice_bytes=8192)
    _, rep = gpu_codec.encode_block(bad, opts)
    assert rep.status == 1                                          # load_seqs returns NULL
    block, rep = gpu_codec.encode_block(text, opts)
    dmg = np.array(block, copy=True)
    dmg[len(dmg) // 2] ^= 0x10
    back, drep = gpu_codec.decode_block(dmg, len(text) + 64)
    assert back is None and drep.status == 3                        # "Block CRC mismatch" (fqzcomp5.c:2312-2317)
    back, drep = gpu_codec.decode_block(block[:40], len(text) + 64)
    assert back is None and drep.status == 1


def test_block_caller_name_coder(gpu_codec, checker):
    """n_name_methods = 0: the caller's coder (here: a stand-in for tok3) turns the name buffer into the
    name section on the host while the device runs the trials; its bytes land in the block verbatim."""
    from oracle.pyoracle import FastqChecker, fastq_available
    fqc = FastqChecker("ref" if fastq_available("ref") else "oracle")
    text = fastq_text(400, 150, seed=8, pair=True)
    want = fqc.split(text)
    seen = {}
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def coder(user, names, name_len, flags, nrec, out, out_len):
        raw = C.string_at(names, name_len)
        seen["names"], seen["flags"] = raw, [flags[i] for i in range(nrec)]
        sec = struct.pack("<IBI", name_len, 0x7f, name_len) + raw          # [u32 ulen][strat][u32 clen][payload]
        p = libc.malloc(len(sec))
        C.memmove(p, sec, len(sec))
        out[0] = p
        out_len[0] = len(sec)
        return 0
    cb = gpu_codec.NAME_CODER(coder)
    opts = gpu_codec.block_opts(slice_bytes=65536, names=(), name_coder=cb)
    block, rep = gpu_codec.encode_block(text, opts)
    assert rep.status == 0
    assert seen["names"] == want["name"] and seen["flags"] == want["flag"]
    b = bytes(block)
    assert b[12:12 + 9 + len(want["name"])] == struct.pack("<IBI", len(want["name"]), 0x7f, len(want["name"])) + want["name"]
    assert struct.unpack_from("<I", b, 8)[0] == zlib.crc32(b[12:])


def test_blocks_multi_in_dispatch_order(gpu_codec, checker):
    """b200fqz_encode_blocks_multi / decode_blocks_multi on the persistent workers (two per device):
    block b's result is in slot b whatever finished first (thread_pool.c:113-164)."""
    import torch
    ngpu = min(torch.cuda.device_count(), 2)
    texts = [fastq_text(300 + 211 * i, 100 + 10 * (i % 3), seed=20 + i) for i in range(7)]
    arrs = [np.frombuffer(t, np.uint8) for t in texts]
    outs = [np.empty(int(gpu_codec.lib().b200fqz_block_bound(a.size)), np.uint8) for a in arrs]
    opts = gpu_codec.block_opts(slice_bytes=30000, x32=True)
    reps = gpu_codec.encode_blocks_multi(ngpu, arrs, opts, outs)
    single = [gpu_codec.encode_block(t, opts)[0].tobytes() for t in texts]
    for b in range(7):
        assert reps[b].status == 0
        assert outs[b][:reps[b].block_len].tobytes() == single[b], b
    backs = [np.empty(a.size + 64, np.uint8) for a in arrs]
    dreps = gpu_codec.decode_blocks_multi(ngpu, outs, [r.block_len for r in reps], backs)
    for b in range(7):
        assert dreps[b].status == 0 and backs[b][:dreps[b].block_len].tobytes() == texts[b]
