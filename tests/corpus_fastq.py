"""FASTQ block texts for the split / join parity tests (deterministic)."""
import numpy as np


def illumina(nrec, read_len=150, seed=1, paired=False, plus_name=False):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(nrec):
        name = b"SIM.%d %d/%d" % (i // 2 if paired else i, i, (i % 2) + 1 if paired else 1)
        seq = rng.choice(np.frombuffer(b"ACGTN", np.uint8), read_len, p=[.25, .25, .25, .24, .01]).tobytes()
        qual = (rng.integers(2, 41, read_len) + 33).astype(np.uint8).tobytes()
        out.append(b"@" + name + b"\n" + seq + b"\n+" + (name if plus_name else b"") + b"\n" + qual + b"\n")
    return b"".join(out)


def long_reads(nrec, seed=4, lo=200, hi=9000):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(nrec):
        n = int(rng.integers(lo, hi))
        seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes()
        qual = (rng.integers(1, 41, n) + 33).astype(np.uint8).tobytes()
        out.append(b"@read_%d ch=%d\n" % (i, i % 512) + seq + b"\n+\n" + qual + b"\n")
    return b"".join(out)


def edge_cases():
    """(label, text).  Every case is fed to the reference, the oracle and the GPU."""
    base = (b"@r1/1\nACGT\n+\nIIII\n@r1/2\nACGTA\n+r1\nIIIII\n@dup\nAC\n+\n##\n@dup\nGG\n+\n!!\n"
            b"@/2\nA\n+\nI\n@x/2\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n")
    cases = [("empty", b""), ("base", base), ("first_is_slash2", b"@/2\nAC\n+\nII\n@/2\nAC\n+\nII\n"),
             ("empty_fields", b"@\n\n+\n\n@\n\n+\n\n"), ("only_newlines", b"\n\n\n\n"),
             ("no_at", b"r1\nAC\n+\nII\n"), ("no_at_second", b"@r1\nAC\n+\nII\nr2\nAC\n+\nII\n"),
             ("no_plus", b"@r1\nAC\n-\nII\n"), ("len_mismatch_mid", b"@r1\nACG\n+\nII\n@r2\nAC\n+\nII\n"),
             ("len_mismatch_last_nl", b"@r1\nAC\n+\nII\n@r2\nACG\n+\nII\n"),
             ("len_mismatch_last_no_nl", b"@r1\nAC\n+\nII\n@r2\nACG\n+\nII"),
             ("no_trailing_newline", b"@r1\nAC\n+\nII\n@r2\nAC\n+\nII"),
             ("partial_bad_at", b"@r1\nAC\n+\nII\nXr2\nAC"), ("partial_bad_plus", b"@r1\nAC\n+\nII\n@r2\nAC\n-"),
             ("partial_seq_nl_last", b"@r1\nAC\n+\nII\n@r2\nACG\n"),
             ("partial_seq_then_plus", b"@r1\nAC\n+\nII\n@r2\nACG\n+"),
             ("partial_changes_fixed_len", b"@r1\nAC\n+\nII\n@r2\nACG\n+\nI"),
             ("fixed_zero_first", b"@a\n\n+\n\n@b\nAC\n+\nII\n"),
             ("crlf", b"@r1\r\nAC\r\n+\r\nII\r\n"),
             ("high_bytes", b"@r\xff\nAC\n+\n\xff\x80\n")]
    for cut in range(len(base) + 1):          # every prefix: all the end-of-block rules
        cases.append(("prefix%d" % cut, base[:cut]))
    il = illumina(40, 37, seed=3, paired=True)
    for cut in (len(il), len(il) - 1, len(il) - 38, len(il) - 39, len(il) - 40, len(il) // 2):
        cases.append(("illumina_cut%d" % cut, il[:cut]))
    cases.append(("illumina_plus_name", illumina(25, 50, seed=5, plus_name=True)))
    cases.append(("long_reads", long_reads(12, seed=6)))
    cases.append(("illumina_8k_tile_edges", illumina(300, 150, seed=7)))
    return cases
