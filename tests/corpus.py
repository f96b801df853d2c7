"""Deterministic inputs shared by the oracle tests, the GPU parity tests and the
golden-vector generator (SURVEY 8c's differential corpus)."""
import numpy as np

from fqzcomp5_b200 import synth

ORDERS_BASE = [0, 1, 64, 65, 128, 129, 192, 193]
SIZES_EDGE = [0, 1, 7, 8, 19, 20, 21, 31, 32, 33, 999, 1000, 1001, 4097]
SIZES_MID = [49999, 50000, 300000]


def make(name, n, seed=1):
    """bytes of length n."""
    if n == 0:
        return b""
    rng = np.random.default_rng(seed)
    if name in synth.GENERATORS:
        return synth.GENERATORS[name](n, seed=seed).tobytes()
    if name == "random":                      # incompressible -> CAT
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if name == "const":                       # one symbol: freq 4096
        return b"A" * n
    if name.startswith("nsym"):               # exactly k symbols, skewed: pack boundaries 2/3/4/5/16/17
        k = int(name[4:])
        p = np.arange(1, k + 1, dtype=np.float64) ** -1.2
        return (rng.choice(k, n, p=p / p.sum()).astype(np.uint8) * 7 + 3).tobytes()
    if name == "runs":                        # long runs: RLE accepted
        v = rng.integers(0, 6, max(1, n // 20 + 1), dtype=np.uint8)
        ln = rng.integers(1, 60, v.size)
        return np.repeat(v, ln)[:n].tobytes().ljust(n, b"\0")
    if name == "noruns":                      # alternating: RLE rejected
        return (np.arange(n) % 3).astype(np.uint8).tobytes()
    if name == "text":                        # many symbols, order-1 structure, big o1 table
        words = [b"@SIM.", b"read/", b"ACGTN", b" len=", b"0123456789", b"\n", b"flowcell:lane:tile:"]
        out = bytearray()
        i = 0
        while len(out) < n:
            out += words[int(rng.integers(0, len(words)))] + str(i).encode()
            i += 1
        return bytes(out[:n])
    if name == "wide":                        # all 256 symbols, mild skew: o1 table > 1000 B (self-compressed)
        p = np.arange(1, 257, dtype=np.float64) ** -0.7
        return rng.choice(256, n, p=p / p.sum()).astype(np.uint8).tobytes()
    if name == "stripe32":                    # 4-byte little-endian integers (STRIPE's purpose)
        v = np.cumsum(rng.integers(0, 50, n // 4 + 1)).astype("<u4")
        return v.tobytes()[:n]
    raise KeyError(name)


GENS = ["illumina_qual", "illumina_seq", "ont_qual", "binned_qual", "random", "const", "nsym2", "nsym3",
        "nsym4", "nsym5", "nsym16", "nsym17", "runs", "noruns", "text", "wide", "stripe32"]


def orders_all():
    o = []
    for b in ORDERS_BASE:
        o += [b, b | 4]
    o += [0x20, 0x10, 0x14, 0x11, 0x15, 0xd5]                      # CAT, NOSZ
    o += [(N << 8) | 8 for N in (0, 2, 4, 150)]                    # STRIPE o0
    o += [(N << 8) | 9 for N in (0, 2, 4, 150)]                    # STRIPE o0/o1
    o += [(4 << 8) | 0xcd, (4 << 8) | 0xc9 | (1 << 16)]            # STRIPE with all methods / NO0
    o += [1 << 17, (1 << 17) | 1, (1 << 17) | 0xc1]                # SIMD_AUTO
    return o


def parity_cases(sizes, gens=None, orders=None):
    for g in (gens or GENS):
        for n in sizes:
            for o in (orders or orders_all()):
                yield g, n, 1, o


def golden_cases():
    """Small but covering: every generator x a few sizes x the main orders."""
    cases = []
    for g in GENS:
        for n in (33, 1001, 70000):
            for o in (0, 1, 4, 5, 0xc0, 0xc1, 0xc5, (4 << 8) | 9):
                cases.append((g, n, 1, o))
    for n in SIZES_EDGE:
        for o in (0, 5, 0xc5, 0x15, 0x408):
            cases.append(("illumina_qual", n, 2, o))
    cases.append(("illumina_qual", 1 << 20, 2, 4))
    cases.append(("illumina_qual", 1 << 20, 2, 5))
    cases.append(("ont_qual", 1 << 20, 4, 5))
    cases.append(("illumina_seq", 1 << 20, 3, 0xc5))
    cases.append(("illumina_qual", 600000, 2, (150 << 8) | 9))
    return cases
