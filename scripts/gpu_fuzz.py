"""Randomised differential test: GPU library vs the reference build, both directions.
usage: gpu_fuzz.py [cases] [seed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fqzcomp5_b200 import codec
from oracle.pyoracle import Codec, available
ref = Codec("ref") if available("ref") else Codec("oracle")
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)

def gen(n):
    kind = rng.integers(0, 8)
    if kind == 0:   # skewed alphabet of random size
        k = int(rng.integers(1, 257)); p = rng.random(k) ** rng.uniform(1, 6); p /= p.sum()
        syms = rng.permutation(256)[:k].astype(np.uint8)
        return syms[rng.choice(k, n, p=p)]
    if kind == 1:   # markov / sticky
        k = int(rng.integers(2, 60)); stay = rng.uniform(0.3, 0.98)
        v = rng.integers(0, k, n).astype(np.uint8); keep = rng.random(n) < stay; keep[0] = False
        idx = np.where(~keep, np.arange(n), 0); np.maximum.accumulate(idx, out=idx)
        return (v[idx] + rng.integers(0, 200)).astype(np.uint8)
    if kind == 2:   # few symbols (pack)
        k = int(rng.integers(1, 18)); return (rng.integers(0, k, n) * int(rng.integers(1, 14))).astype(np.uint8)
    if kind == 3:   # long runs (rle)
        m = max(1, n // int(rng.integers(3, 200))); v = rng.integers(0, int(rng.integers(2, 30)), m + 1).astype(np.uint8)
        ln = rng.integers(1, max(2, 2 * n // m), m + 1); return np.repeat(v, ln)[:n] if np.repeat(v, ln).size >= n else np.resize(np.repeat(v, ln), n)
    if kind == 4:   # uniform random
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == 5:   # little-endian integers (stripe)
        w = int(rng.choice([2, 4, 8])); x = np.cumsum(rng.integers(0, 1000, n // w + 1)).astype("<u%d" % w)
        return np.frombuffer(x.tobytes()[:n], np.uint8)
    if kind == 6:   # one dominant symbol, rare others (tiny frequencies, normalisation corner cases)
        a = np.full(n, int(rng.integers(0, 256)), np.uint8); m = rng.random(n) < rng.uniform(0.0001, 0.02)
        a[m] = rng.integers(0, 256, int(m.sum())); return a
    return (np.arange(n) * int(rng.integers(1, 9)) % int(rng.integers(2, 256))).astype(np.uint8)

base_orders = [0, 1, 4, 5, 0x40, 0x41, 0x44, 0x45, 0x80, 0x81, 0x84, 0x85, 0xc0, 0xc1, 0xc4, 0xc5]
bad = 0
for it in range(ncase):
    n = int(rng.choice([rng.integers(0, 64), rng.integers(64, 2000), rng.integers(2000, 70000), rng.integers(70000, 600000)]))
    d = gen(n).tobytes() if n else b""
    assert len(d) == n
    o = int(rng.choice(base_orders))
    r = rng.random()
    if r < 0.15: o = (int(rng.choice([0, 2, 3, 4, 7, 150])) << 8) | 8 | (o & 0xc5)
    elif r < 0.2: o |= 0x10
    elif r < 0.25: o |= 1 << 17
    want = ref.compress(d, o)
    got = codec.rans_compress_to_4x16(d, o)
    if want != got:
        bad += 1; print("ENC", it, n, hex(o), want and (len(want), hex(want[0])), got and (len(got), hex(got[0])))
        continue
    if want is not None:
        back = codec.rans_uncompress_to_4x16(want, n) if want[0] & 0x10 else codec.rans_uncompress_4x16(want)
        if back != d:
            bad += 1; print("DEC", it, n, hex(o), hex(want[0]))
print("cases", ncase, "mismatches", bad)
