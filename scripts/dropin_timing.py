"""Whole-tool timing: stock fqzcomp5 objects linked against libb200rans.so vs the reference binary."""
import os, subprocess, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fqzcomp5_b200 import synth
OBJ = os.path.join(ROOT, "oracle", "_ref", "fqzobj")
tmp = tempfile.mkdtemp()
exe = os.path.join(tmp, "fqz_gpu")
objs = [os.path.join(OBJ, f) for f in sorted(os.listdir(OBJ)) if f.endswith(".o")]
libdir = os.path.join(ROOT, "fqzcomp5_b200")
subprocess.run(["gcc", "-o", exe] + objs + ["-L" + libdir, "-lb200rans", "-Wl,-rpath," + libdir, "-lz", "-lm", "-pthread"], check=True)
ref = os.path.join(ROOT, "oracle", "_ref", "fqzcomp5_ref")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
seq = synth.illumina_seq(n * 150).reshape(n, 150)
qual = (synth.illumina_qual(n * 150) + 33).astype(np.uint8).reshape(n, 150)
fq = os.path.join(tmp, "in.fq")
with open(fq, "wb") as f:
    for i in range(n):
        f.write(b"@SIM.%d %d/1\n" % (i, i) + seq[i].tobytes() + b"\n+\n" + qual[i].tobytes() + b"\n")
print("fastq bytes", os.path.getsize(fq))
for level in ("-1", "-3"):
    for name, binary in (("reference", ref), ("gpu drop-in", exe)):
        for t in ("1", "8"):
            out = os.path.join(tmp, "o.fqz5")
            t0 = time.time(); subprocess.run([binary, level, "-t", t, fq, out], check=True, capture_output=True); t1 = time.time()
            subprocess.run([binary, "-d", "-t", t, out, os.path.join(tmp, "b.fq")], check=True, capture_output=True); t2 = time.time()
            print("%-12s %s -t %s: compress %.2f s, decompress %.2f s, size %d" % (name, level, t, t1 - t0, t2 - t1, os.path.getsize(out)))
