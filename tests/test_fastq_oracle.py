"""The FASTQ split / join oracle (oracle/fastq_oracle.c) against the unmodified reference's
load_seqs() / output_fastq() and against the committed vectors generated from them."""
import hashlib
import json
import os
import sys

import pytest

import corpus_fastq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_fastq import describe  # noqa: E402


@pytest.fixture(scope="module")
def fq_oracle():
    from oracle.pyoracle import FastqChecker
    return FastqChecker("oracle")


@pytest.fixture(scope="module")
def fq_ref():
    from oracle.pyoracle import FastqChecker, fastq_available, REFERENCE_ROOT
    if not fastq_available("ref") and not os.path.isdir(REFERENCE_ROOT):
        pytest.skip("oracle/_ref/libref_fqz.so not built and /root/reference absent")
    return FastqChecker("ref")


@pytest.fixture(scope="module")
def fq_golden():
    with open(os.path.join(ROOT, "tests", "golden", "fastq_vectors.json")) as f:
        return {v["label"]: v for v in json.load(f)["vectors"]}


def test_oracle_matches_golden_vectors(fq_oracle, fq_golden):
    cases = corpus_fastq.edge_cases()
    assert len(cases) == len(fq_golden)
    for label, text in cases:
        r = fq_oracle.split(text)
        v = describe(label, text, r)
        if r is not None:
            v["join0_sha256"] = hashlib.sha256(fq_oracle.join(r["name"], r["seq"], r["qual"], r["len"], 0)).hexdigest()
            v["join1_sha256"] = hashlib.sha256(fq_oracle.join(r["name"], r["seq"], r["qual"], r["len"], 1)).hexdigest()
        assert v == fq_golden[label], label


def test_oracle_matches_reference(fq_oracle, fq_ref):
    for label, text in corpus_fastq.edge_cases() + [("big", corpus_fastq.illumina(3000, 150, seed=9, paired=True)),
                                                     ("ont", corpus_fastq.long_reads(40, seed=10))]:
        a, b = fq_oracle.split(text), fq_ref.split(text)
        assert a == b, label
        if a is not None:
            for p in (0, 1):
                assert fq_oracle.join(a["name"], a["seq"], a["qual"], a["len"], p) == \
                    fq_ref.join(a["name"], a["seq"], a["qual"], a["len"], p), label


def test_join_inverts_split(fq_oracle):
    text = corpus_fastq.illumina(500, 100, seed=12)
    r = fq_oracle.split(text)
    assert r["consumed"] == len(text) and r["fixed_len"] == 100
    assert fq_oracle.join(r["name"], r["seq"], r["qual"], r["len"], 0) == text


def test_oracle_matches_reference_on_random_texts(fq_oracle, fq_ref):
    """Random record soups, cut anywhere, damaged in one byte, with NUL bytes (which load_seqs reads as
    line ends -- the one point where the GPU path deliberately differs and reports a malformed block)."""
    import numpy as np
    rng = np.random.default_rng(77)
    n_null = 0
    for it in range(400):
        recs = []
        for i in range(int(rng.integers(0, 60))):
            n = int(rng.integers(0, 120))
            nm = bytes(rng.integers(33, 127, int(rng.integers(0, 30))).astype(np.uint8))
            if recs and rng.random() < 0.2:
                nm = recs[-1][0]
            if rng.random() < 0.2:
                nm += b"/2"
            recs.append((nm, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n)),
                         bytes((rng.integers(0, 60, n) + 33).astype(np.uint8))))
        t = b"".join(b"@" + a + b"\n" + s + b"\n+\n" + q + b"\n" for a, s, q in recs)
        r = rng.random()
        if r < 0.4 and t:
            t = t[:int(rng.integers(0, len(t) + 1))]
        elif r < 0.7 and len(t) > 4:
            b = bytearray(t)
            b[int(rng.integers(0, len(b)))] = int(rng.choice([0, 10, 43, 64, 65]))
            t = bytes(b)
        a, b = fq_oracle.split(t), fq_ref.split(t)
        assert a == b, (it, t[:80])
        n_null += a is None
    assert 0 < n_null < 400


def _same_blocks(a, b):
    if a is None or b is None:
        return a is None and b is None
    keys = ("num_records", "name", "seq", "qual", "len", "flag", "fixed_len")
    return len(a) == len(b) and all(all(x[k] == y[k] for k in keys) for x, y in zip(a, b))


def test_kseq_oracle_matches_reference(fq_oracle, fq_ref):
    """fqo_split_kseq against the unmodified load_seqs_kseq (fqzcomp5.c:423-623, the loader main() reaches),
    called block after block on the same file as encode_gzip does (:3051)."""
    for label, text, blk in corpus_fastq.kseq_cases() + corpus_fastq.kseq_bad_cases():
        a, b = fq_oracle.split_kseq(text, blk), fq_ref.split_kseq(text, blk)
        assert _same_blocks(a, b), (label, a if a is None else len(a), b if b is None else len(b))
        if a:
            assert sum(x["consumed"] for x in a) == len(text), label      # every byte of a well-formed file is taken
