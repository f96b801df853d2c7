"""Parity of the CUDA path with the CPU checkers, through the C ABI.

Bit-exact both ways: the GPU's compressed bytes equal the CPU
codec's, and the GPU decodes CPU-encoded streams back to the input.  The
checker is the unmodified reference build when oracle/_ref is present, else
the oracle; the committed golden vectors pin both.
"""
import hashlib

import numpy as np
import pytest

import corpus

pytestmark = pytest.mark.gpu


def _gpu_decode(g, comp, n):
    return g.rans_uncompress_to_4x16(comp, n) if comp[0] & 0x10 else g.rans_uncompress_4x16(comp)


def _cpu_decode(c, comp, n):
    return c.uncompress(comp, n) if comp[0] & 0x10 else c.uncompress(comp)


@pytest.mark.parametrize("gen", corpus.GENS)
def test_single_call_parity(gpu_codec, checker, gen):
    """Every order flag x edge sizes, one rans_compress_to_4x16 call each."""
    bad = []
    for g, n, seed, order in corpus.parity_cases(corpus.SIZES_EDGE + [50000], gens=[gen]):
        data = corpus.make(g, n, seed)
        want = checker.compress(data, order)
        got = gpu_codec.rans_compress_to_4x16(data, order)
        if want != got:
            bad.append(("enc", g, n, hex(order), want and len(want), got and len(got)))
            continue
        if want is not None and _gpu_decode(gpu_codec, want, n) != data:
            bad.append(("dec", g, n, hex(order)))
    assert not bad, bad[:10]


def test_golden_vectors(gpu_codec, golden):
    """The reference's own outputs, committed under tests/golden/."""
    bad = []
    for v in golden["vectors"]:
        data = corpus.make(v["gen"], v["n"], v["seed"])
        out = gpu_codec.rans_compress_to_4x16(data, v["order"])
        if v.get("null"):
            ok = out is None
        else:
            ok = (out is not None and len(out) == v["len"]
                  and hashlib.sha256(out).hexdigest() == v["sha256"]
                  and _gpu_decode(gpu_codec, out, len(data)) == data)
        if not ok:
            bad.append((v["gen"], v["n"], hex(v["order"])))
    assert not bad, bad[:10]


def test_malloc_forms(gpu_codec, checker):
    data = corpus.make("ont_qual", 70000, 4)
    for order in (0, 5, 0xc5, (4 << 8) | 9):
        want = checker.compress_malloc(data, order)
        assert gpu_codec.rans_compress_4x16(data, order) == want
        assert gpu_codec.rans_uncompress_4x16(want) == data


def test_capacity_semantics(gpu_codec, checker):
    """out != NULL: *out_size is the capacity; too small => NULL, like the reference."""
    data = corpus.make("illumina_qual", 20000, 2)
    for order in (0, 5, 0x45):
        full = checker.bound(len(data), order)
        for cap in (1, 10, 1000, full - 100, full):
            assert gpu_codec.rans_compress_to_4x16(data, order, cap=cap) == checker.compress(data, order, cap=cap), \
                (hex(order), cap)


def test_mid_sizes_and_self_compressed_tables(gpu_codec, checker):
    cases = [("wide", 300000, 5), ("wide", 300000, 1), ("text", 300000, 5), ("text", 49999, 1),
             ("illumina_qual", 500001, 5), ("illumina_qual", 500001, 4), ("ont_qual", 1 << 20, 5),
             ("illumina_seq", 1 << 20, 0xc5), ("binned_qual", 1 << 20, 0x84), ("runs", 300000, 0x45),
             ("illumina_qual", 600000, (150 << 8) | 9), ("stripe32", 400000, (4 << 8) | 9),
             ("nsym17", 300000, 0xc5), ("nsym16", 300000, 0xc1)]
    for g, n, order in cases:
        data = corpus.make(g, n, 1)
        want = checker.compress(data, order)
        assert gpu_codec.rans_compress_to_4x16(data, order) == want, (g, n, hex(order))
        assert _gpu_decode(gpu_codec, want, n) == data, (g, n, hex(order))


def test_corrupt_streams_fail_cleanly(gpu_codec, checker):
    """Truncated / damaged input: NULL or garbage of the right length, never a fault."""
    data = corpus.make("illumina_qual", 6000, 2)
    for order in (0, 1, 4, 5, 0xc5, 0x408):
        c = checker.compress(data, order)
        assert gpu_codec.rans_uncompress_4x16(c[:10]) is None or order == 0x408
        assert gpu_codec.rans_uncompress_4x16(b"") is None
        rng = np.random.default_rng(5)
        for _ in range(8):
            d = bytearray(c)
            pos = int(rng.integers(1, len(d)))
            d[pos] ^= 1 << int(rng.integers(0, 8))
            r = gpu_codec.rans_uncompress_4x16(bytes(d))
            assert r is None or isinstance(r, bytes)
        # the library is still healthy afterwards
        assert gpu_codec.rans_uncompress_4x16(c) == data


def test_batch_host_api(gpu_codec, checker):
    """Many streams per call, mixed orders and sizes, one arena out."""
    parts, orders = [], []
    for g, n, o in [("illumina_qual", 262144, 4), ("illumina_qual", 262144, 5), ("ont_qual", 100000, 5),
                    ("illumina_seq", 200000, 0xc5), ("random", 5000, 4), ("const", 70000, 0x84),
                    ("text", 30000, 1), ("illumina_qual", 0, 0), ("runs", 90000, 0x44),
                    ("illumina_qual", 60000, (150 << 8) | 9), ("nsym4", 777, 0xc1)] * 3:
        parts.append(np.frombuffer(corpus.make(g, n, 3), np.uint8))
        orders.append(o)
    sizes = [p.size for p in parts]
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    buf = np.concatenate(parts) if sum(sizes) else np.zeros(0, np.uint8)
    out, ooff, osz = gpu_codec.compress_batch(buf, offs, sizes, orders)
    for k, (p, o) in enumerate(zip(parts, orders)):
        want = checker.compress(p.tobytes(), o)
        got = out[int(ooff[k]):int(ooff[k]) + int(osz[k])].tobytes()
        assert got == want, (k, hex(o), len(got), len(want))
    back = np.empty(buf.size + 16, np.uint8)
    rsz, status = gpu_codec.uncompress_batch(out, ooff, osz, back, offs, sizes)
    assert (status == 0).all() and (rsz == np.array(sizes, np.uint32)).all()
    assert np.array_equal(back[:buf.size], buf)


def test_batch_device_api_roundtrip(gpu_codec, checker):
    """Device-resident path (what bench.py's `value` times): parity per slice."""
    import torch
    data = np.frombuffer(corpus.make("illumina_qual", 4 << 20, 2), np.uint8)
    for order in (4, 5):
        S = 262144
        sl = [(o, min(S, data.size - o)) for o in range(0, data.size, S)]
        in_off = np.array([o for o, _ in sl], np.uint64)
        in_size = np.array([s for _, s in sl], np.uint32)
        orders = np.full(len(sl), order, np.int32)
        d_in = torch.from_numpy(data.copy()).cuda()
        cap = gpu_codec.compress_bound_batch(in_size, orders)
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        d_off = torch.zeros(len(sl), dtype=torch.int64, device="cuda")
        d_sz = torch.zeros(len(sl), dtype=torch.int32, device="cuda")
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            gpu_codec.compress_batch_dev(st.cuda_stream, d_in.data_ptr(), in_off, in_size, orders,
                                         d_out.data_ptr(), cap, d_off.data_ptr(), d_sz.data_ptr())
        st.synchronize()
        off, sz, comp = d_off.cpu().numpy(), d_sz.cpu().numpy(), d_out.cpu().numpy()
        for k, (o, s) in enumerate(sl):
            assert comp[off[k]:off[k] + sz[k]].tobytes() == checker.compress(data[o:o + s].tobytes(), order), k
        d_back = torch.zeros(data.size, dtype=torch.uint8, device="cuda")
        d_osz = torch.zeros(len(sl), dtype=torch.int32, device="cuda")
        d_st = torch.ones(len(sl), dtype=torch.int32, device="cuda")
        flags = comp[off]
        with torch.cuda.stream(st):
            gpu_codec.uncompress_batch_dev(st.cuda_stream, d_out.data_ptr(), off.astype(np.uint64),
                                           sz.astype(np.uint32), d_back.data_ptr(), in_off, in_size,
                                           d_osz.data_ptr(), d_st.data_ptr(), flags=flags)
        st.synchronize()
        assert int(d_st.abs().sum()) == 0
        assert torch.equal(d_back, d_in)


def test_large_block_roundtrip_property(gpu_codec):
    """Size-independent property at a block the oracle would take minutes on:
    encode -> decode is the identity, and the total compressed size matches the
    sum over a sample of slices checked individually elsewhere."""
    n = 256 << 20
    from fqzcomp5_b200 import synth
    buf = synth.illumina_qual(n)
    S = 262144
    sl = synth.slices(buf, S)
    offs = [o for o, _ in sl]
    sizes = [s for _, s in sl]
    for order in (4, 5):
        out, ooff, osz = gpu_codec.compress_batch(buf, offs, sizes, [order] * len(sl))
        assert (osz > 0).all()
        back = np.empty(n, np.uint8)
        rsz, status = gpu_codec.uncompress_batch(out, ooff, osz, back, offs, sizes)
        assert (status == 0).all()
        assert hashlib.sha256(back).digest() == hashlib.sha256(buf).digest()
        assert 0.05 < float(osz.sum()) / n < 0.8


def test_concurrent_callers(gpu_codec, checker):
    """hts_tpool calls the codec from several worker threads at once (SURVEY 8b threading)."""
    import threading
    data = [corpus.make("illumina_qual", 50000 + 1000 * i, i) for i in range(8)]
    want = [checker.compress(d, 5) for d in data]
    res = [None] * 8

    def work(i):
        for _ in range(3):
            res[i] = gpu_codec.rans_compress_to_4x16(data[i], 5)
            assert gpu_codec.rans_uncompress_4x16(res[i]) == data[i]
    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert res == want


def test_unaligned_buffers_and_odd_word_streams(gpu_codec, checker):
    """Streams, inputs and outputs at every byte alignment: the coders' vector paths and the
    word ring must not depend on alignment (the word stream itself starts at an odd address
    for about half of all tables)."""
    base = np.frombuffer(corpus.make("illumina_qual", 200000, 7), np.uint8)
    for shift in range(0, 17, 3):
        for order in (4, 5):
            n = 150000 + shift
            buf = np.zeros(n + 64, np.uint8)
            buf[shift:shift + n] = base[:n]
            out, ooff, osz = gpu_codec.compress_batch(buf, [shift], [n], [order])
            want = checker.compress(buf[shift:shift + n].tobytes(), order)
            got = out[int(ooff[0]):int(ooff[0]) + int(osz[0])].tobytes()
            assert got == want, (shift, order)
            # decode from an arena where the stream sits at an odd offset, into an odd offset
            comp = np.zeros(len(want) + 64, np.uint8)
            comp[shift + 1:shift + 1 + len(want)] = np.frombuffer(want, np.uint8)
            back = np.zeros(n + 64, np.uint8)
            rsz, st = gpu_codec.uncompress_batch(comp, [shift + 1], [len(want)], back, [shift + 3], [n])
            assert st[0] == 0 and rsz[0] == n
            assert np.array_equal(back[shift + 3:shift + 3 + n], buf[shift:shift + n]), (shift, order)


def test_one_large_stream_per_call(gpu_codec, checker):
    """A whole buffer as ONE call (K = 1, the single-warp latency floor of SURVEY F4):
    still the reference's bytes."""
    data = corpus.make("illumina_qual", 24 << 20, 11)
    for order in (4, 5, 0, 1):
        want = checker.compress(data, order)
        assert gpu_codec.rans_compress_to_4x16(data, order) == want, hex(order)
        assert gpu_codec.rans_uncompress_4x16(want) == data, hex(order)


def _multi_roundtrip(gpu_codec, checker, ngpu):
    parts = [np.frombuffer(corpus.make("illumina_qual", 262144, s), np.uint8) for s in range(12)]
    buf = np.concatenate(parts)
    offs = [262144 * i for i in range(12)]
    sizes = [262144] * 12
    orders = [4, 5] * 6
    block_of = [i // 3 for i in range(12)]
    for _ in range(2):          # the second call runs on the workers and contexts the first one created
        out, ooff, osz = gpu_codec.compress_batch(buf, offs, sizes, orders, ngpu=ngpu, block_of=block_of, multi=True)
        for k in range(12):
            assert out[int(ooff[k]):int(ooff[k]) + int(osz[k])].tobytes() == checker.compress(parts[k].tobytes(), orders[k])
        back = np.empty(buf.size, np.uint8)
        rsz, st = gpu_codec.uncompress_batch(out, ooff, osz, back, offs, sizes, ngpu=ngpu, block_of=block_of, multi=True)
        assert (st == 0).all() and np.array_equal(back, buf)


def test_multi_batch_api_on_persistent_worker(gpu_codec, checker):
    """b200rans_*_batch_multi with ngpu = 1: the call still runs on the library's persistent worker
    thread (its own context and arenas), results in call order."""
    _multi_roundtrip(gpu_codec, checker, 1)
    with pytest.raises(gpu_codec.B200RansError):
        gpu_codec.compress_batch(np.zeros(8, np.uint8), [0], [8], [0], ngpu=1, block_of=[-1], multi=True)


def test_multi_gpu_batch_api(gpu_codec, checker):
    """b200rans_*_batch_multi on TWO devices: blocks dealt round-robin, one persistent worker and
    context per device, results gathered in call order.  Skipped (not downgraded) on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    _multi_roundtrip(gpu_codec, checker, 2)


def test_decode_without_flags(gpu_codec, checker):
    """b200rans_uncompress_batch_dev with flags == NULL ("unknown"): order-0, order-1, raw and
    PACK / RLE streams all decode through the general kernel."""
    import torch
    items = [("illumina_qual", 100000, 4), ("illumina_qual", 100000, 5), ("illumina_qual", 70000, 0),
             ("ont_qual", 90000, 1), ("illumina_seq", 120000, 0xc5), ("runs", 50000, 0x44), ("random", 3000, 4),
             ("binned_qual", 80000, 0x84), ("text", 60000, 5), ("illumina_qual", 0, 5)]
    raw = [corpus.make(g, n, 3) for g, n, _ in items]
    comp = [checker.compress(d, o) for d, (_, _, o) in zip(raw, items)]
    coff = np.concatenate([[0], np.cumsum([(len(c) + 15) & ~15 for c in comp])[:-1]]).astype(np.uint64)
    arena = np.zeros(int(coff[-1]) + len(comp[-1]) + 64, np.uint8)
    for c, o in zip(comp, coff):
        arena[int(o):int(o) + len(c)] = np.frombuffer(c, np.uint8)
    sizes = np.array([len(d) for d in raw], np.uint32)
    ooff = np.concatenate([[0], np.cumsum((sizes.astype(np.int64) + 15) // 16 * 16)[:-1]]).astype(np.uint64)
    d_in = torch.from_numpy(arena).cuda()
    d_out = torch.zeros(int(ooff[-1]) + int(sizes[-1]) + 64, dtype=torch.uint8, device="cuda")
    d_osz = torch.zeros(len(items), dtype=torch.int32, device="cuda")
    d_st = torch.ones(len(items), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    gpu_codec.uncompress_batch_dev(st.cuda_stream, d_in.data_ptr(), coff, np.array([len(c) for c in comp], np.uint32),
                                   d_out.data_ptr(), ooff, sizes, d_osz.data_ptr(), d_st.data_ptr(), flags=None)
    st.synchronize()
    assert d_st.cpu().tolist() == [0] * len(items)
    back = d_out.cpu().numpy()
    for d, o in zip(raw, ooff):
        assert back[int(o):int(o) + len(d)].tobytes() == d


def test_in_slot_output(gpu_codec, checker):
    """b200rans_compress_batch_dev2 with OUT_IN_SLOT: every stream stays in its own bound-sized slot
    (no packing pass) and is still the reference's bytes -- coded, raw (CAT), STRIPE, tiny, empty,
    self-compressed order-1 tables, PACK / RLE."""
    import torch
    items = [("illumina_qual", 262144, 4), ("illumina_qual", 262144, 5), ("ont_qual", 100001, 5),
             ("illumina_seq", 200000, 0xc5), ("random", 5000, 4), ("const", 70000, 0x84), ("text", 30000, 1),
             ("illumina_qual", 0, 0), ("runs", 90000, 0x44), ("illumina_qual", 60000, (150 << 8) | 9),
             ("nsym4", 777, 0xc1), ("wide", 300000, 5), ("illumina_qual", 19, 5), ("stripe32", 40000, (4 << 8) | 9),
             ("illumina_qual", 33, 4), ("text", 300000, 5)] * 2
    parts = [np.frombuffer(corpus.make(g, n, 3), np.uint8) for g, n, _ in items]
    orders = np.array([o for _, _, o in items], np.int32)
    sizes = np.array([p.size for p in parts], np.uint32)
    offs = np.concatenate([[0], np.cumsum((sizes.astype(np.int64) + 15) // 16 * 16)[:-1]]).astype(np.uint64)
    buf = np.zeros(int(offs[-1]) + int(sizes[-1]) + 64, np.uint8)
    for p_, o in zip(parts, offs):
        buf[int(o):int(o) + p_.size] = p_
    d_in = torch.from_numpy(buf).cuda()
    cap = gpu_codec.compress_slots_bound(sizes, orders)
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(len(items), dtype=torch.int64, device="cuda")
    d_sz = torch.zeros(len(items), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    gpu_codec.compress_batch_dev2(st.cuda_stream, d_in.data_ptr(), offs, sizes, orders, d_out.data_ptr(), cap,
                                  d_off.data_ptr(), d_sz.data_ptr(), flags=gpu_codec.OUT_IN_SLOT)
    st.synchronize()
    off, sz, comp = d_off.cpu().numpy(), d_sz.cpu().numpy(), d_out.cpu().numpy()
    for k, (p_, o) in enumerate(zip(parts, orders)):
        want = checker.compress(p_.tobytes(), int(o))
        assert comp[off[k]:off[k] + sz[k]].tobytes() == want, (k, items[k], int(sz[k]), len(want))
    # and the slots decode where they lie
    d_back = torch.zeros(buf.size, dtype=torch.uint8, device="cuda")
    d_osz = torch.zeros(len(items), dtype=torch.int32, device="cuda")
    d_st = torch.ones(len(items), dtype=torch.int32, device="cuda")
    plain = [k for k in range(len(items)) if not (comp[off[k]] & 8)]         # STRIPE goes through the host-buffer API
    gpu_codec.uncompress_batch_dev(st.cuda_stream, d_out.data_ptr(), off[plain].astype(np.uint64),
                                   sz[plain].astype(np.uint32), d_back.data_ptr(), offs[plain], sizes[plain],
                                   d_osz.data_ptr(), d_st.data_ptr(), flags=comp[off[plain]])
    st.synchronize()
    assert int(d_st[:len(plain)].abs().sum()) == 0
    back = d_back.cpu().numpy()
    for k in plain:
        assert np.array_equal(back[int(offs[k]):int(offs[k]) + int(sizes[k])], parts[k]), k


def test_histogram_pass_alignments_and_ranges(gpu_codec, checker):
    """hist_kernel (counts and pair counts in front of the coders): inputs at odd device addresses, sizes around
    its loop boundaries, and symbol ranges either side of the limit of its byte-indexed pair matrix (a span of 90
    symbols fits, 91 goes through the rank-space form), with and without symbol 0 in the data, sticky and not."""
    import torch
    rng = np.random.default_rng(11)

    def sticky(n, lo, hi, keep):
        draw = rng.random(n) >= keep
        draw[0] = True
        vals = rng.integers(lo, hi + 1, n, dtype=np.uint8)
        idx = np.where(draw, np.arange(n), 0)
        np.maximum.accumulate(idx, out=idx)
        return vals[idx]

    items = []
    for n in (4096, 4097, 4111, 32767, 32768, 32769 + 16, 65536 + 15, 262144 + 7):
        for lo, hi, keep in ((2, 40, 0.875), (0, 40, 0.875), (10, 99, 0.5), (10, 100, 0.5), (1, 200, 0.9), (0, 255, 0.0),
                             (7, 7, 0.0), (33, 71, 0.0)):
            items.append((sticky(n, lo, hi, keep), 5 if (n + hi) % 3 else 1))
    items += [(sticky(70001, 2, 40, 0.875), 4), (sticky(5000, 2, 40, 0.875), 0), (np.repeat(np.arange(2, 41, dtype=np.uint8), 997), 5)]
    sizes = np.array([d.size for d, _ in items], np.uint32)
    orders = np.array([o for _, o in items], np.int32)
    offs, at = [], 0
    for k, (d, _) in enumerate(items):
        at += k % 16                       # every alignment of the first byte
        offs.append(at)
        at += d.size
    offs = np.array(offs, np.uint64)
    buf = np.zeros(at + 64, np.uint8)
    for (d, _), o in zip(items, offs):
        buf[int(o):int(o) + d.size] = d
    d_in = torch.from_numpy(buf).cuda()
    cap = gpu_codec.compress_slots_bound(sizes, orders)
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(len(items), dtype=torch.int64, device="cuda")
    d_sz = torch.zeros(len(items), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    gpu_codec.compress_batch_dev2(st.cuda_stream, d_in.data_ptr(), offs, sizes, orders, d_out.data_ptr(), cap,
                                  d_off.data_ptr(), d_sz.data_ptr(), flags=gpu_codec.OUT_IN_SLOT)
    st.synchronize()
    off, sz, comp = d_off.cpu().numpy(), d_sz.cpu().numpy(), d_out.cpu().numpy()
    for k, (d, o) in enumerate(items):
        want = checker.compress(d.tobytes(), int(o))
        assert comp[off[k]:off[k] + sz[k]].tobytes() == want, (k, d.size, int(d.min()), int(d.max()), int(o))


def test_method_trial_device_resident(gpu_codec, checker):
    """b200rans_compress_trials_dev: ragged trial with inputs and winners in HBM, winners packed without
    gaps (pack_align 1) as the block pipeline uses it."""
    import torch
    lists = [[0, 1, 129, 193, gpu_codec.ransxn1_order(150)], [4, 5, 133, 197], [1], [193, 0]]
    items = [("illumina_qual", 150 * 300, 1), ("illumina_seq", 150 * 900, 2), ("runs", 70000, 3), ("random", 9000, 4),
             ("binned_qual", 150 * 500, 5), ("illumina_qual", 0, 6), ("text", 50000, 7), ("const", 20000, 8)]
    parts = [np.frombuffer(corpus.make(g, n, s), np.uint8) for g, n, s in items]
    ml = [lists[i % len(lists)] for i in range(len(items))]
    sizes = np.array([p_.size for p_ in parts], np.uint32)
    offs = np.concatenate([[0], np.cumsum((sizes.astype(np.int64) + 15) // 16 * 16)[:-1]]).astype(np.uint64)
    buf = np.zeros(int(offs[-1]) + int(sizes[-1]) + 64, np.uint8)
    for p_, o in zip(parts, offs):
        buf[int(o):int(o) + p_.size] = p_
    d_in = torch.from_numpy(buf).cuda()
    cap = int(sum(max(gpu_codec.rans_compress_bound_4x16(int(s), m) for m in l) for s, l in zip(sizes, ml))) + 4096
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    n, mm = len(items), sum(len(l) for l in ml)
    d_off = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_sz = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_best = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    d_cs = torch.zeros(mm, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    first = gpu_codec.compress_trials_dev(st.cuda_stream, d_in.data_ptr(), offs, sizes, ml, d_out.data_ptr(), cap, 1,
                                          d_off.data_ptr(), d_sz.data_ptr(), d_best.data_ptr(), d_cs.data_ptr())
    st.synchronize()
    off, sz, best, cs, comp = (t.cpu().numpy() for t in (d_off, d_sz, d_best, d_cs, d_out))
    at = 0
    for k in range(n):
        wb, wout, wsizes = _cpu_trial(checker, parts[k].tobytes(), ml[k])
        assert cs[first[k]:first[k + 1]].tolist() == wsizes and best[k] == wb, (k, items[k])
        assert off[k] == at, "winners are packed without gaps, in input order"
        assert comp[off[k]:off[k] + sz[k]].tobytes() == wout
        at += int(sz[k])


def test_randomised_differential(gpu_codec, checker):
    """A slice of scripts/gpu_fuzz.py: random distributions, sizes and order flags."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "gpu_fuzz.py"), "400", "99"],
                         capture_output=True, text=True, timeout=600).stdout
    assert "cases 400 mismatches 0" in out, out[-2000:]


def test_precision_decision_at_its_boundary(gpu_codec, checker):
    """rans_compute_shift (rANS_static4x16pr.c:357-420) is taken on the device in double precision; inputs whose
    e10 / e12 sits within 1e-3 of the 1.01 threshold (found by bisection on the length, scripts/gpu_fuzz_shift.py)
    must still give the reference's bytes."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "gpu_fuzz_shift.py"), "6", "7"],
                         capture_output=True, text=True, timeout=900).stdout
    assert "mismatches 0" in out and "boundary cases 0 " not in out, out[-2000:]


# ---------------------------------------------------------------- method trial
def _cpu_trial(checker, data, methods):
    """compress_with_methods restated on the CPU checker (fqzcomp5.c:1989-2106, rANS members):
    try the methods in list order, keep the first strictly smaller stream."""
    best, best_out, sizes = -1, None, []
    for j, order in enumerate(methods):
        out = checker.compress_malloc(data, order)
        sizes.append(len(out) if out else 0)
        if out and (best_out is None or len(best_out) > len(out)):
            best, best_out = j, out
    return best, best_out, sizes


def test_method_trial_batch(gpu_codec, checker):
    """fqzcomp5 -3 / -5 method sets for the seq and qual sections, many inputs per call."""
    qual_methods = [0, 1, 129, 193, gpu_codec.ransxn1_order(150)]         # fqzcomp5.c:4893-4900 + RANSXN1
    seq_methods = [0, 1, 64, 65, 128, 129, 192, 193]                      # fqzcomp5.c:2005
    cases = [
        (qual_methods, [("illumina_qual", 150 * 400, 1), ("binned_qual", 150 * 700, 2), ("ont_qual", 90000, 3),
                        ("const", 30000, 1), ("runs", 60000, 4), ("illumina_qual", 150 * 7, 5),
                        ("illumina_qual", 0, 1), ("random", 40000, 6), ("nsym4", 19, 1)]),
        (seq_methods, [("illumina_seq", 200000, 1), ("nsym4", 50000, 2), ("nsym16", 50000, 3),
                       ("nsym17", 50000, 4), ("text", 70000, 5), ("wide", 120000, 6), ("random", 1000, 7)]),
    ]
    for methods, items in cases:
        parts = [np.frombuffer(corpus.make(g, n, s), np.uint8) for g, n, s in items]
        sizes = [p.size for p in parts]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        buf = np.concatenate(parts)
        out, ooff, osz, best, csize = gpu_codec.compress_methods_batch(buf, offs, sizes, methods)
        for k, (g, n, s) in enumerate(items):
            wb, wout, wsizes = _cpu_trial(checker, parts[k].tobytes(), methods)
            assert list(csize[k]) == wsizes, (g, n, list(csize[k]), wsizes)
            assert best[k] == wb, (g, n, best[k], wb)
            got = out[int(ooff[k]):int(ooff[k]) + int(osz[k])].tobytes()
            assert got == wout, (g, n)
            assert gpu_codec.rans_uncompress_4x16(got) == parts[k].tobytes()


def test_method_trial_single(gpu_codec, checker):
    data = corpus.make("illumina_qual", 150 * 2000, 9)
    for methods in ([0, 1, 129, 193], [1], [193, 129, 1, 0], [gpu_codec.ransxn1_order(150), 1]):
        wb, wout, wsizes = _cpu_trial(checker, data, methods)
        got, best, csize = gpu_codec.compress_methods(data, methods)
        assert (got, best, list(csize)) == (wout, wb, wsizes), methods


def test_method_trial_many_inputs(gpu_codec, checker):
    """More inputs than one pipeline chunk holds; winners differ between inputs."""
    gens = ["illumina_qual", "binned_qual", "runs", "random", "ont_qual", "const"]
    methods = [4, 5, 0x84, 0x45, 0xc5]
    parts = [np.frombuffer(corpus.make(gens[i % len(gens)], 20000 + 997 * (i % 13), i), np.uint8) for i in range(700)]
    sizes = [p.size for p in parts]
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    buf = np.concatenate(parts)
    out, ooff, osz, best, csize = gpu_codec.compress_methods_batch(buf, offs, sizes, methods)
    assert len(set(best.tolist())) > 1
    for k in range(0, 700, 7):
        wb, wout, wsizes = _cpu_trial(checker, parts[k].tobytes(), methods)
        assert best[k] == wb and list(csize[k]) == wsizes
        assert out[int(ooff[k]):int(ooff[k]) + int(osz[k])].tobytes() == wout


def test_tok3_stream_trials(gpu_codec, checker):
    """tok3's compress() (tokenise_name3.c:1268-1417) as one ragged trial: every token stream of a
    block brute-forced over the method list of its type and level; 4-lane streams, STRIPE N=4,
    PACK and RLE candidates, first smallest kept."""
    rng = np.random.default_rng(11)
    # token-stream-like payloads: type bytes, small deltas, little-endian 32-bit digits, text
    def payload(t, n):
        if t in (0, 10, 11, 12):
            return rng.choice(np.array([1, 7, 8, 9, 10], np.uint8), n, p=[.05, .6, .2, .1, .05])
        if t in (3, 5, 6, 7):
            v = (rng.integers(0, 50000, n // 4 + 1).astype("<u4") + np.arange(n // 4 + 1, dtype="<u4") * 3)
            return v.view(np.uint8)[:n]
        if t in (8, 9, 4):
            return rng.choice(np.array([0, 1, 2, 3, 255], np.uint8), n, p=[.1, .7, .1, .05, .05])
        return np.frombuffer(corpus.make("text", n, int(rng.integers(1, 99))), np.uint8)
    for level in (1, 3, 5, 7, 9):
        row = gpu_codec.TOK3_METHODS[gpu_codec.tok3_level_row(level)]
        parts, lists = [], []
        for t in range(13):
            for n in (0, 3, 40, 1000, 4001, 24000):
                if n == 0 and t:
                    continue
                lst = gpu_codec.tok3_method_list(row[t], n)
                if not lst:
                    continue
                parts.append(np.ascontiguousarray(payload(t, n)))
                lists.append(lst)
        sizes = [p.size for p in parts]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        buf = np.concatenate(parts)
        out, ooff, osz, best, cs = gpu_codec.compress_trials(buf, offs, sizes, lists)
        for k in range(len(parts)):
            wb, wout, wsizes = _cpu_trial(checker, parts[k].tobytes(), lists[k])
            assert cs[k] == wsizes, (level, k, lists[k], cs[k], wsizes)
            assert best[k] == wb
            assert out[int(ooff[k]):int(ooff[k]) + int(osz[k])].tobytes() == wout


def test_staged_decode_large_alphabets(gpu_codec, checker):
    """Order-1 streams behind PACK / RLE whose (packed) alphabet has more than 64 symbols take the staged decode
    (head / table / chain / post launches, dec_staged.cu); smaller ones are handed back to the general kernel.
    Reference-coded streams must decode to the input, 4 and 32 lanes, every transform combination."""
    rng = np.random.default_rng(77)
    def sticky(n, k, stay=0.75):                     # k symbols, each repeating its predecessor with p = stay: the
        v = rng.integers(0, k, n).astype(np.uint8)    # packed bytes use all 256 values and stay compressible
        keep = rng.random(n) < stay
        keep[0] = False
        idx = np.where(~keep, np.arange(n), 0)
        np.maximum.accumulate(idx, out=idx)
        return (v[idx] * 3 + 65).astype(np.uint8).tobytes()
    def walk(n, k=200):                               # a random walk over 200 symbols in short runs: RLE alone, large alphabet
        m = n // 3 + 2
        q = np.cumsum(rng.integers(-2, 3, m)) % k
        return np.repeat(q.astype(np.uint8), rng.integers(1, 7, m))[:n].tobytes().ljust(n, b"\1")
    cases = []
    for n in (1500, 4096, 30000, 70001, 262144, 300007):
        cases += [(sticky(n, 4), o) for o in (0xc5, 0xc1, 0x85, 0x81, 0xd5)]
        cases += [(sticky(n, 16), o) for o in (0xc5, 0x81)]
        cases += [(sticky(n, 2), 0xc5), (sticky(n, 5), 0xc1), (sticky(n, 4, 0.97), 0xc5), (sticky(n, 16, 0.97), 0xc1)]
        cases += [(walk(n), o) for o in (0x45, 0x41)]
    cases += [(corpus.make("illumina_seq", 1 << 20, 3), 0xc5), (corpus.make("wide", 200000, 3), 0x45)]
    bad = []
    gpu_codec.dec_staged_stats(reset=True)
    for data, order in cases:
        want = checker.compress(data, order)
        assert want is not None
        if _gpu_decode(gpu_codec, want, len(data)) != data:
            bad.append(("dec", len(data), hex(order), hex(want[0])))
        if gpu_codec.rans_compress_to_4x16(data, order) != want:
            bad.append(("enc", len(data), hex(order)))
    assert not bad, bad[:10]
    stats = gpu_codec.dec_staged_stats()
    assert stats[0] >= len(cases) // 3 and stats[1] == 0, stats     # the staged route really took them
    # a batch: many staged streams next to plain ones in one call
    parts = [np.frombuffer(d, np.uint8) for d, _ in cases[:24]]
    orders = [o for _, o in cases[:24]]
    sizes = [p.size for p in parts]
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    buf = np.concatenate(parts)
    out, ooff, osz = gpu_codec.compress_batch(buf, offs, sizes, orders)
    back = np.zeros(buf.size, np.uint8)
    dsz, st = gpu_codec.uncompress_batch(out, ooff, osz, back, offs, sizes)
    assert all(int(x) == 0 for x in st) and [int(x) for x in dsz] == sizes
    assert np.array_equal(back, buf)
    # damaged staged streams: a result or a clean failure, and the library stays healthy
    data, order = cases[0]
    c = checker.compress(sticky(50000, 4, 0.97), 0xc5)
    for _ in range(40):
        d = bytearray(c)
        pos = int(rng.integers(1, len(d)))
        d[pos] ^= 1 << int(rng.integers(0, 8))
        r = gpu_codec.rans_uncompress_4x16(bytes(d[:int(rng.integers(pos, len(d) + 1))]))
        assert r is None or isinstance(r, bytes)
    assert gpu_codec.rans_uncompress_4x16(checker.compress(data, order)) == data
