"""Parity of the FASTQ split / join kernels (fqzcomp5_b200/csrc/fastq.cu) with the reference's
load_seqs() / output_fastq() (oracle/_ref/libref_fqz.so when present, else the oracle), and with
the committed vectors generated from the reference."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import corpus_fastq

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def fq_checker():
    from oracle.pyoracle import FastqChecker, fastq_available
    return FastqChecker("ref" if fastq_available("ref") else "oracle")


def test_split_join_golden_vectors(gpu_codec):
    from make_golden_fastq import describe
    with open(os.path.join(ROOT, "tests", "golden", "fastq_vectors.json")) as f:
        gold = {v["label"]: v for v in json.load(f)["vectors"]}
    bad = []
    for label, text in corpus_fastq.edge_cases():
        r = gpu_codec.load_seqs(text)
        v = describe(label, text, r)
        if r is not None:
            for p in (0, 1):
                v["join%d_sha256" % p] = hashlib.sha256(
                    gpu_codec.output_fastq(r["name"], r["seq"], r["qual"], r["len"], p)).hexdigest()
        if v != gold[label]:
            bad.append(label)
    assert not bad, bad[:20]


def test_split_matches_checker_on_blocks(gpu_codec, fq_checker):
    """Larger blocks: many 8 KiB tiles, block ends inside every field, paired names, long reads."""
    texts = [corpus_fastq.illumina(20000, 150, seed=21, paired=True),
             corpus_fastq.illumina(5000, 151, seed=22, plus_name=True),
             corpus_fastq.long_reads(300, seed=23, lo=100, hi=40000)]
    for ti, text in enumerate(texts):
        for cut in (len(text), len(text) - 1, len(text) - 77, len(text) - 200, (len(text) * 2) // 3):
            blk = text[:cut]
            want = fq_checker.split(blk)
            got = gpu_codec.load_seqs(blk)
            assert got == want, (ti, cut)
            for p in (0, 1):
                assert gpu_codec.output_fastq(got["name"], got["seq"], got["qual"], got["len"], p) == \
                    fq_checker.join(want["name"], want["seq"], want["qual"], want["len"], p), (ti, cut, p)
            assert gpu_codec.output_fastq(got["name"], got["seq"], got["qual"], got["len"], 0) == \
                blk[:got["consumed"]] or ti == 1


def test_split_random_lengths(gpu_codec, fq_checker):
    rng = np.random.default_rng(31)
    for it in range(20):
        recs = []
        for i in range(int(rng.integers(1, 400))):
            n = int(rng.integers(0, 300))
            nm = bytes(rng.integers(33, 127, int(rng.integers(0, 40))).astype(np.uint8)).replace(b"\n", b"_")
            if rng.random() < 0.2 and recs:
                nm = recs[-1][0]
            if rng.random() < 0.2:
                nm += b"/2"
            recs.append((nm, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n)),
                         bytes((rng.integers(0, 60, n) + 33).astype(np.uint8))))
        text = b"".join(b"@" + a + b"\n" + s + b"\n+\n" + q + b"\n" for a, s, q in recs)
        cut = int(rng.integers(0, len(text) + 1)) if it % 2 else len(text)
        assert gpu_codec.load_seqs(text[:cut]) == fq_checker.split(text[:cut]), it


def test_split_capacity_and_nul(gpu_codec):
    text = corpus_fastq.illumina(100, 50, seed=41)
    with pytest.raises(gpu_codec.B200RansError):
        gpu_codec.load_seqs(text, max_records=10)
    # documented deviation: a NUL byte makes the block malformed (the reference reads it as a line end)
    assert gpu_codec.load_seqs(text[:200] + b"\0" + text[200:]) is None


def test_device_resident_split_feeds_codec(gpu_codec, checker):
    """FASTQ text in HBM -> split on the device -> the seq / qual buffers go straight into
    b200rans_compress_batch_dev; the streams equal the CPU codec's on the CPU-split buffers."""
    import ctypes as C
    import torch
    from oracle.pyoracle import FastqChecker
    text = corpus_fastq.illumina(30000, 150, seed=51)
    want = FastqChecker("oracle").split(text)
    n = len(text)
    dev = torch.device("cuda", 0)
    d_text = torch.from_numpy(np.frombuffer(text, np.uint8).copy()).to(dev)
    mr = n // 100
    d_name = torch.empty(n, dtype=torch.uint8, device=dev); d_seq = torch.empty(n, dtype=torch.uint8, device=dev)
    d_qual = torch.empty(n, dtype=torch.uint8, device=dev)
    d_len = torch.empty(mr, dtype=torch.int32, device=dev); d_flag = torch.empty(mr, dtype=torch.int32, device=dev)
    d_no = torch.empty(mr, dtype=torch.int32, device=dev); d_so = torch.empty(mr, dtype=torch.int32, device=dev)
    L = gpu_codec.lib()
    sb = int(L.b200fq_split_scratch_bytes(n, mr))
    d_scr = torch.empty(sb + 256, dtype=torch.uint8, device=dev)
    scr = (d_scr.data_ptr() + 255) & ~255
    d_info = torch.zeros(16, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.b200fq_split_dev(st, d_text.data_ptr(), n, d_name.data_ptr(), n, d_seq.data_ptr(), d_qual.data_ptr(), n,
                            d_len.data_ptr(), d_flag.data_ptr(), d_no.data_ptr(), d_so.data_ptr(), mr, scr, sb,
                            d_info.data_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    info = d_info.cpu().numpy()
    assert info[0] == 0 and info[1] == want["num_records"] and info[3] == len(want["seq"])
    assert d_qual[:info[4]].cpu().numpy().tobytes() == want["qual"]
    assert d_so[:info[1]].cpu().numpy().tolist() == list(np.cumsum([0] + want["len"][:-1]))
    # qualities straight into the codec: 1000 reads per call
    S = 150 * 1000
    k = info[4] // S
    in_off = np.arange(k, dtype=np.uint64) * S
    in_size = np.full(k, S, np.uint32)
    orders = np.full(k, 5, np.int32)
    cap = gpu_codec.compress_bound_batch(in_size, orders)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(k, dtype=torch.int64, device=dev); d_sz = torch.zeros(k, dtype=torch.int32, device=dev)
    gpu_codec.compress_batch_dev(st, d_qual.data_ptr(), in_off, in_size, orders, d_out.data_ptr(), cap,
                                 d_off.data_ptr(), d_sz.data_ptr())
    torch.cuda.synchronize()
    off, sz = d_off.cpu().numpy(), d_sz.cpu().numpy()
    out = d_out.cpu().numpy()
    for j in (0, k // 2, k - 1):
        assert out[off[j]:off[j] + sz[j]].tobytes() == checker.compress(want["qual"][j * S:(j + 1) * S], 5)


def _gpu_kseq_blocks(gpu_codec, text, blk):
    """load_seqs_kseq's blocks of a text, one GPU call per block, each fed from where the last one stopped
    (and only as much text as a caller reading ahead would have: a little more than the block needs)."""
    pos, out = 0, []
    while True:
        r = gpu_codec.load_seqs_kseq(text[pos:], blk)
        if r is None:
            return None
        if r["num_records"] == 0:
            break
        out.append(r)
        pos += r["consumed"]
        if not r["more"]:
            break
    return out


def test_kseq_mode_matches_live_loader(gpu_codec, fq_checker):
    """B200FQ_MODE_KSEQ against load_seqs_kseq (fqzcomp5.c:423-623), the loader main() really reaches:
    kseq's name / comment split and re-join, the block-size rule with its buffered record, READ2, fixed_len."""
    keys = ("num_records", "name", "seq", "qual", "len", "flag", "fixed_len")
    big = corpus_fastq.illumina(30000, 150, seed=31, paired=True)
    cases = corpus_fastq.kseq_cases() + corpus_fastq.kseq_bad_cases() + [("kbig", big, 1 << 20), ("kbig_all", big, 1 << 30)]
    for label, text, blk in cases:
        want = fq_checker.split_kseq(text, blk)
        got = _gpu_kseq_blocks(gpu_codec, text, blk)
        if want is None or got is None:
            assert want is None and got is None, label
            continue
        assert len(got) == len(want), (label, len(got), len(want))
        for i, (g, w) in enumerate(zip(got, want)):
            for k in keys:
                assert g[k] == w[k], (label, i, k)
        assert sum(g["consumed"] for g in got) == len(text), label


def test_block_call_in_kseq_mode(gpu_codec, fq_checker):
    """b200fqz_encode_block with kseq_blk_size: a file walked block by block as encode_gzip does (:3051-3077),
    every block decoding back to the text it consumed (names without comments: join restores the header)."""
    text = corpus_fastq.illumina(4000, 100, seed=41, paired=True)
    blk = 150000
    want = fq_checker.split_kseq(text, blk)
    opts = gpu_codec.block_opts(slice_bytes=65536, x32=True)
    opts.kseq_blk_size = blk
    pos = 0
    for i, w in enumerate(want):
        block, rep = gpu_codec.encode_block(text[pos:pos + 2 * blk + 4096], opts)
        assert rep.status == 0 and rep.num_records == w["num_records"], i
        assert (rep.ulen[0], rep.ulen[1], rep.ulen[2]) == (len(w["name"]), len(w["seq"]), len(w["qual"])), i
        back, drep = gpu_codec.decode_block(block, 3 * blk)
        assert drep.status == 0 and back.tobytes() == text[pos:pos + rep.consumed], i
        pos += rep.consumed
    assert pos == len(text)
