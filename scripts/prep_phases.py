"""Per-phase cycle counts of prep_kernel (needs a library built with -DB200_PREP_PROF; see scripts/build_debug.sh)."""
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from fqzcomp5_b200 import codec, synth
n = 954
d = synth.illumina_seq(n * 262144, seed=3)
sizes = [262144] * n
offs = (np.arange(n) * 262144).astype(np.uint64)
for it in range(2):
    out, ooff, osz = codec.compress_batch(d, offs, sizes, [0xc5] * n)
o = (C.c_ulonglong * 16)()
codec.lib().b200rans_prep_prof(o)
v = [int(x) for x in o]
tot = sum(v) or 1
names = ["pack", "rle", "hist8", "-", "alphabet+buckets", "sweep1", "sweep2", "table pack", "E stream"]
for i, nm in enumerate(names):
    print("%-18s %6.1f%%  %8.0f cycles per stream" % (nm, 100 * v[i] / tot, v[i] / (2 * n)))
