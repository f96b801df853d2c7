"""Exploratory GPU probe: parity of single calls vs the CPU checkers + quick batch timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fqzcomp5_b200 import codec, synth
from oracle.pyoracle import Codec, available

ref = Codec("ref") if available("ref") else Codec("oracle")
print("checker:", ref.kind, codec.lib().b200rans_version())

def first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], np.uint8); y = np.frombuffer(b[:n], np.uint8)
    d = np.nonzero(x != y)[0]
    return int(d[0]) if d.size else n

orders = [int(x, 0) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4, 1, 5]
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 7, 8, 31, 32, 33, 999, 1000, 1001, 4097, 70000, 300000]
bad = 0
rng = np.random.default_rng(1)
for gen in ["illumina_qual", "illumina_seq", "ont_qual", "binned_qual", "random", "const"]:
    for n in sizes:
        if gen == "random": d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif gen == "const": d = b"A" * n
        else: d = synth.GENERATORS[gen](n).tobytes() if n else b""
        for o in orders:
            want = ref.compress(d, o)
            got = codec.rans_compress_to_4x16(d, o)
            if want != got:
                bad += 1
                print("ENC MISMATCH", gen, n, hex(o), "want", None if want is None else (len(want), hex(want[0])),
                      "got", None if got is None else (len(got), hex(got[0])),
                      "first diff", first_diff(want, got) if want and got else None)
                if want and got:
                    i = first_diff(want, got); print("   want", want[max(0,i-4):i+12].hex(), "got", got[max(0,i-4):i+12].hex())
            if want is not None:
                ulen = len(d)
                back = codec.rans_uncompress_to_4x16(want, ulen) if (want[0] & 0x10) else codec.rans_uncompress_4x16(want)
                if back != d:
                    bad += 1
                    print("DEC MISMATCH", gen, n, hex(o), "flag", hex(want[0]), "got", None if back is None else len(back),
                          "first diff", first_diff(back, d) if back else None)
print("mismatches:", bad)
