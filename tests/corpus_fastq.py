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


def kseq_cases():
    """(label, text, blk_size) for the live loader's rules (load_seqs_kseq, fqzcomp5.c:423-623): strict
    4-line FASTQ ending in a newline -- comments behind spaces and tabs, empty comments, /2 pairs, repeated
    names, fixed and variable lengths, block sizes that cut after one record, mid-stream and never."""
    rng = np.random.default_rng(12)
    base = (b"@r1/1 first comment\nACGT\n+\nIIII\n@r1/2\tcomment after a tab\nACGTA\n+r1\nIIIII\n"
            b"@dup\nAC\n+\n##\n@dup\nGG\n+\n!!\n@dup \nGG\n+\n!!\n@dup x\nGG\n+\n!!\n@dup\tx\nGG\n+\n!!\n"
            b"@/2\nA\n+\nI\n@x/2\nAAAAAAAA\n+\nIIIIIIII\n@y c/2\nAC\n+\nII\n@2 /2\nAC\n+\nII\n"
            b"@a  two spaces\nAC\n+\nII\n@ lead\nAC\n+\nII\n@\n\n+\n\n@e\n\n+\n\n@tab\t\nAC\n+\nII\n")
    cases = [("kbase_all", base, 1 << 20)]
    for blk in (1, 7, 12, 13, 14, 30, 31, 64, 200):
        cases.append(("kbase_blk%d" % blk, base, blk))
    il = illumina(400, 50, seed=21, paired=True)
    for blk in (1000, 4096, 10000, len(il), 10 * len(il)):
        cases.append(("kil_blk%d" % blk, il, blk))
    var = []
    for i in range(300):
        n = int(rng.integers(1, 90))
        var.append(b"@v%d/%d len=%d\n" % (i // 2, i % 2 + 1, n) + b"ACGT"[i % 4:i % 4 + 1] * n + b"\n+\n" +
                   bytes(rng.integers(35, 70, n).astype(np.uint8)) + b"\n")
    var = b"".join(var)
    for blk in (333, 5000):
        cases.append(("kvar_blk%d" % blk, var, blk))
    cases.append(("klong", long_reads(15, seed=8), 30000))
    cases.append(("kempty", b"", 100))
    return cases


def kseq_bad_cases():
    """Texts the reference's kseq loader fails on as well (the strict reading reports status 1)."""
    return [("kbad_qual_long", b"@r1\nAC\n+\nIII\n@r2\nAC\n+\nII\n", 1000),
            ("kbad_qual_long_next_block", b"@r1\nAC\n+\nII\n@r2\nAC\n+\nIII\n", 3)]
