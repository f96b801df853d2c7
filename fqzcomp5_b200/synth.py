"""Deterministic synthetic FASTQ streams for the configs in BASELINE.json.

Shapes follow SURVEY.md 8(d).  All generators are numpy, seeded, and produce the
byte streams fqzcomp5 hands to the codec (quality values already have 33
subtracted, fqzcomp5.c:563-564; sequence is plain ACGT text).  Large outputs are
built from independently seeded 32 MiB chunks so generation parallelises and a
prefix of a large stream equals the same-seed smaller one.
"""
from concurrent.futures import ThreadPoolExecutor
import numpy as np

CHUNK = 1 << 25


def _chunks(n, fn, seed, align=1):
    step = CHUNK - CHUNK % align
    parts = [(i, min(step, n - i)) for i in range(0, n, step)]
    if len(parts) <= 1:
        return fn(n, seed) if n else np.zeros(0, np.uint8)
    with ThreadPoolExecutor(8) as ex:
        out = list(ex.map(lambda p: fn(p[1], seed + 1000003 * (p[0] // step)), parts))
    return np.concatenate(out)


def _illumina_qual(n, seed):
    rng = np.random.default_rng(seed)
    draw = rng.random(n) < 0.125          # keep previous Q with p = 7/8
    draw[0] = True
    vals = rng.integers(2, 41, n, dtype=np.uint8)
    idx = np.where(draw, np.arange(n, dtype=np.int64), 0)
    np.maximum.accumulate(idx, out=idx)
    return vals[idx]


def illumina_qual(n, seed=2):
    """Config 2: 150 bp Illumina quality stream, Q in [2,40], sticky (p=7/8)."""
    return _chunks(n, _illumina_qual, seed)


def binned_qual(n, seed=22):
    """Config 2 variant: NovaSeq-like 4-level binned quals, i.i.d."""
    rng = np.random.default_rng(seed)
    lv = np.array([2, 12, 23, 37], np.uint8)
    return lv[rng.choice(4, n, p=[.02, .05, .13, .8])]


def _illumina_seq(n, seed, read_len=150):
    rng = np.random.default_rng(seed)
    T = np.random.default_rng(3).dirichlet([0.6] * 4, size=64)   # fixed order-3 model
    cum = np.cumsum(T, axis=1)
    nreads = (n + read_len - 1) // read_len
    out = np.empty((nreads, read_len), np.uint8)
    ctx = rng.integers(0, 64, nreads)
    for p in range(read_len):
        u = rng.random(nreads)
        b = (u[:, None] > cum[ctx, :3]).sum(1).astype(np.int64)
        out[:, p] = b
        ctx = ((ctx << 2) | b) & 63
    tails = np.nonzero(rng.random(nreads) < 0.05)[0]              # poly-G tails
    tl = rng.integers(30, 101, tails.size)
    col = np.arange(read_len)
    mask = col[None, :] >= (read_len - tl)[:, None]
    sub = out[tails]
    sub[mask] = 2
    out[tails] = sub
    return np.frombuffer(b"ACGT", np.uint8)[out].reshape(-1)[:n]


def illumina_seq(n, seed=3):
    """Config 3: ACGT from an order-3 Markov chain, 5% of reads with a poly-G tail."""
    return _chunks(n, _illumina_seq, seed, align=150)


def _ont_qual(n, seed):
    rng = np.random.default_rng(seed)
    lens = []
    tot = 0
    while tot < n:
        l = int(rng.integers(10000, 50001))
        lens.append(l)
        tot += l
    nreads, maxlen = len(lens), max(lens)
    lens = np.array(lens)
    q = np.full(nreads, 20, np.int16)
    out = np.empty((nreads, maxlen), np.uint8)
    steps = rng.choice(np.array([-2, -1, 0, 1, 2], np.int8), size=(maxlen, nreads),
                       p=[.1, .2, .4, .2, .1])
    for p in range(maxlen):
        q = np.clip(q + steps[p], 1, 40)
        out[:, p] = q
    keep = np.arange(maxlen)[None, :] < lens[:, None]
    return out[keep][:n]


def ont_qual(n, seed=4):
    """Config 4: ONT long reads (10-50 kb), quality random walk clamped to [1,40]."""
    return _chunks(n, _ont_qual, seed)


def ont_read_bounds(n, seed=4):
    """Read lengths only (for decomposing config 4 by read groups)."""
    rng = np.random.default_rng(seed)
    return rng.integers(10000, 50001, max(1, n // 30000))


GENERATORS = {
    "illumina_qual": illumina_qual,
    "binned_qual": binned_qual,
    "illumina_seq": illumina_seq,
    "ont_qual": ont_qual,
}


def slices(buf, S):
    """Stream decomposition: K calls of S bytes (last one shorter)."""
    n = len(buf)
    return [(o, min(S, n - o)) for o in range(0, n, S)]
