"""The oracle (oracle/rans_oracle.c) against the pins: the committed golden
vectors produced by the unmodified reference, and -- where oracle/_ref exists --
the reference library itself, byte for byte, both directions."""
import hashlib

import pytest

import corpus


def _decode(c, comp, n):
    return c.uncompress(comp, n) if comp[0] & 0x10 else c.uncompress(comp)


def test_oracle_matches_golden_vectors(oracle, golden):
    bad = []
    for v in golden["vectors"]:
        data = corpus.make(v["gen"], v["n"], v["seed"])
        assert hashlib.sha256(data).hexdigest() == v["in_sha256"], "generator drifted: %r" % (v,)
        out = oracle.compress(data, v["order"])
        if v.get("null"):
            ok = out is None
        else:
            ok = out is not None and len(out) == v["len"] and hashlib.sha256(out).hexdigest() == v["sha256"]
            if ok and "hex" in v:
                ok = out.hex() == v["hex"]
            if ok:
                ok = _decode(oracle, out, len(data)) == data
        if not ok:
            bad.append((v["gen"], v["n"], hex(v["order"])))
    assert not bad, bad[:10]


@pytest.mark.parametrize("gen", corpus.GENS)
def test_oracle_matches_reference_library(oracle, ref, gen):
    for g, n, seed, order in corpus.parity_cases(corpus.SIZES_EDGE + [50000], gens=[gen]):
        data = corpus.make(g, n, seed)
        a, b = ref.compress(data, order), oracle.compress(data, order)
        assert a == b, (g, n, hex(order), a and len(a), b and len(b))
        if a is not None:
            assert _decode(oracle, a, n) == data, ("oracle decode of reference stream", g, n, hex(order))
            assert _decode(ref, b, n) == data, ("reference decode of oracle stream", g, n, hex(order))


def test_oracle_malloc_form_and_bound(oracle, ref):
    data = corpus.make("illumina_qual", 5000)
    for order in (0, 5, 0xc5, (4 << 8) | 9):
        assert oracle.compress_malloc(data, order) == ref.compress_malloc(data, order)
    for n in (0, 1, 20, 1000, 65536, 10 ** 6, 10 ** 9, 2 ** 31 - 1):
        for order in (0, 1, 4, 5, 0xc5, 0x45, (150 << 8) | 9, 8):
            assert oracle.bound(n, order) == ref.bound(n, order), (n, hex(order))


def test_survey_known_behaviours(oracle):
    """SURVEY 8c [verified] facts."""
    import numpy as np
    rng = np.random.default_rng(0)
    q = corpus.make("illumina_qual", 900)
    assert oracle.compress(q, 5)[0] == 0x01                      # <=1000 bytes drops X32
    a = oracle.compress(b"A" * 5000, 0x84)
    assert len(a) == 6 and a[0] == 0xA0                          # PACK + CAT, 6 bytes
    e = oracle.compress(b"", 0)
    assert len(e) == 2 and e[0] == 0x20                          # empty -> CAT
    assert oracle.compress(q[:7], 5)[0] == 0x20                  # 7 bytes order 5 -> CAT
    r = rng.integers(0, 256, 100000, dtype=np.uint8).tobytes()
    c = oracle.compress(r, 4)
    assert len(c) == len(r) + 4 and c[0] == 0x24                 # incompressible -> CAT, n+4
    s = oracle.compress(b"A" * 5000, 4)
    assert len(s) == 135                                         # 1 + 2 + 4 + 128, no renorm words
    assert oracle.uncompress(s) == b"A" * 5000


def test_oracle_rejects_corrupt_streams(oracle):
    data = corpus.make("illumina_qual", 4000)
    for order in (0, 1, 4, 5):
        c = bytearray(oracle.compress(data, order))
        assert oracle.uncompress(bytes(c[:10])) is None          # truncated below the states
        c2 = bytes(c[:3]) + bytes([c[3] ^ 0xff]) + bytes(c[4:])  # damage the table
        r = oracle.uncompress(c2)                                # must not crash; result may be None or garbage
        assert r is None or len(r) == len(data)
