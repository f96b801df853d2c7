"""Randomised differential test of the rows around the codec: FASTQ split / join against the reference's
load_seqs / output_fastq (or the oracle), CRC-32 against zlib, method trials against the serial loop.
usage: gpu_fuzz_around.py [cases] [seed]"""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fqzcomp5_b200 import codec
from oracle.pyoracle import Codec, FastqChecker, available, fastq_available

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2468)
fq = FastqChecker("ref" if fastq_available("ref") else "oracle")
ref = Codec("ref") if available("ref") else Codec("oracle")


def fastq_text():
    recs, fixed = [], int(rng.integers(0, 2))
    L = int(rng.integers(0, 400))
    for i in range(int(rng.integers(0, 1500))):
        n = L if fixed else int(rng.integers(0, int(rng.choice([60, 400, 30000]))))
        nm = bytes(rng.integers(33, 127, int(rng.integers(0, 60))).astype(np.uint8))
        if recs and rng.random() < 0.15:
            nm = recs[-1][0]
        if rng.random() < 0.15:
            nm += b"/2"
        plus = nm if rng.random() < 0.1 else b""
        recs.append((nm, bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), n)),
                     bytes((rng.integers(0, 60, n) + 33).astype(np.uint8)), plus))
    t = b"".join(b"@" + a + b"\n" + s + b"\n+" + p + b"\n" + q + b"\n" for a, s, q, p in recs)
    r = rng.random()
    if r < 0.5 and t:
        t = t[:int(rng.integers(0, len(t) + 1))]
    elif r < 0.6 and len(t) > 10:                      # damage: a wrong marker or a wrong length
        b = bytearray(t); b[int(rng.integers(0, len(b)))] = int(rng.integers(33, 127)); t = bytes(b)
    return t


bad = 0
for it in range(ncase):
    t = fastq_text()
    want, got = fq.split(t), codec.load_seqs(t)
    if want != got:
        bad += 1; print("SPLIT", it, len(t)); continue
    if want is not None:
        p = int(rng.integers(0, 2))
        if codec.output_fastq(got["name"], got["seq"], got["qual"], got["len"], p) != \
                fq.join(want["name"], want["seq"], want["qual"], want["len"], p):
            bad += 1; print("JOIN", it, len(t))
    n = int(rng.choice([rng.integers(0, 600), rng.integers(600, 300000), rng.integers(300000, 3000000)]))
    b = rng.integers(0, 256, n).astype(np.uint8).tobytes()
    c0 = int(rng.integers(0, 2 ** 32))
    if codec.crc32(b, c0) != zlib.crc32(b, c0):
        bad += 1; print("CRC", it, n)
    if it % 4 == 0:                                    # a small ragged trial
        parts = [rng.integers(0, int(rng.integers(2, 60)), int(rng.integers(0, 40000))).astype(np.uint8) for _ in range(5)]
        lists = [[int(x) for x in rng.choice([0, 1, 4, 5, 64, 65, 128, 129, 192, 193, 0x408, 0x409], int(rng.integers(1, 5)))]
                 for _ in parts]
        sizes = [p.size for p in parts]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        out, ooff, osz, best, cs = codec.compress_trials(np.concatenate(parts), offs, sizes, lists)
        for k, p in enumerate(parts):
            w = [ref.compress_malloc(p.tobytes(), o) for o in lists[k]]
            ws = [len(x) if x else 0 for x in w]
            wb = min((i for i in range(len(w)) if w[i]), key=lambda i: (ws[i], i), default=-1)
            if cs[k] != ws or best[k] != wb or (wb >= 0 and out[int(ooff[k]):int(ooff[k]) + int(osz[k])].tobytes() != w[wb]):
                bad += 1; print("TRIAL", it, k, lists[k], cs[k], ws)
print("cases", ncase, "mismatches", bad)
