"""Generate tests/golden/rans_vectors.json from the UNMODIFIED reference library.

Run in the build container (needs /root/reference to have been compiled by
oracle/Makefile into oracle/_ref/libref_rans.so):

    python tests/golden/make_golden.py

Each vector records how to rebuild the input (generator, size, seed), the order
argument, and the reference's output: the full compressed stream in hex when it
is small, otherwise its length, first 32 bytes and SHA-256.  Also records the
known answer of `fqzcomp5 -1 sample.fastq` (SURVEY 8c / BASELINE.md 2).
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.pyoracle import Codec  # noqa: E402
import corpus  # noqa: E402


def main():
    ref = Codec("ref")
    simd = Codec("ref_simd")
    vectors = []
    for name, n, seed, order in corpus.golden_cases():
        data = corpus.make(name, n, seed)
        out = ref.compress(data, order)
        assert out == simd.compress(data, order), (name, n, hex(order))
        v = {"gen": name, "n": n, "seed": seed, "order": order,
             "in_sha256": hashlib.sha256(data).hexdigest()}
        if out is None:
            v["null"] = True
        else:
            v.update({"len": len(out), "flag": out[0], "sha256": hashlib.sha256(out).hexdigest(),
                      "head": out[:32].hex()})
            if len(out) <= 600:
                v["hex"] = out.hex()
            # the reference must decode its own stream
            back = ref.uncompress(out, len(data)) if out[0] & 0x10 else ref.uncompress(out)
            assert back == data
        vectors.append(v)
    doc = {"source": "oracle/_ref/libref_rans.so built from /root/reference by oracle/Makefile",
           "vectors": vectors}
    exe = os.path.join(ROOT, "oracle", "_ref", "fqzcomp5_ref")
    if os.path.exists(exe):
        fq = os.path.join(HERE, "sample.fastq")
        for lvl in ("-1", "-3"):
            tmp = "/tmp/_golden_sample.fqz5"
            subprocess.run([exe, lvl, fq, tmp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            b = open(tmp, "rb").read()
            doc.setdefault("fqzcomp5", {})[lvl] = {"len": len(b), "md5": hashlib.md5(b).hexdigest(),
                                                   "hex": b.hex()}
    json.dump(doc, open(os.path.join(HERE, "rans_vectors.json"), "w"), indent=0)
    print("wrote %d vectors" % len(vectors))


if __name__ == "__main__":
    main()
