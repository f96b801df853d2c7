"""Show where GPU and checker streams differ for given (gen, n, order) cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import corpus
from fqzcomp5_b200 import codec
from oracle.pyoracle import Codec, available
ref = Codec("ref") if available("ref") else Codec("oracle")
cases = [c.split(":") for c in sys.argv[1:]]
for g, n, o in cases:
    n, o = int(n), int(o, 0)
    d = corpus.make(g, n, 1)
    a, b = ref.compress(d, o), codec.rans_compress_to_4x16(d, o)
    print(g, n, hex(o), "ref", None if a is None else len(a), "gpu", None if b is None else len(b))
    if a and b:
        m = min(len(a), len(b))
        x, y = np.frombuffer(a[:m], np.uint8), np.frombuffer(b[:m], np.uint8)
        df = np.nonzero(x != y)[0]
        print("  ndiff", df.size, "first", df[:8].tolist(), "head ref", a[:12].hex(), "gpu", b[:12].hex())
        for i in df[:3]:
            print("   @%d ref %s gpu %s" % (i, a[max(0, i - 6):i + 10].hex(), b[max(0, i - 6):i + 10].hex()))
