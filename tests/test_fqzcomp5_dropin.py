"""Config 1 (BASELINE.json configs[0]): stock fqzcomp5 objects linked against
libb200rans.so instead of the reference's rANS Nx16 objects must reproduce the
reference's file byte for byte and round-trip it."""
import hashlib
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "oracle", "_ref", "fqzobj")


@pytest.fixture(scope="module")
def fqz_gpu(tmp_path_factory, gpu_codec):
    if not os.path.exists(os.path.join(OBJ, ".done")):
        pytest.skip("stock fqzcomp5 objects (oracle/_ref/fqzobj) were not built")
    exe = str(tmp_path_factory.mktemp("fqz") / "fqzcomp5_b200")
    objs = [os.path.join(OBJ, f) for f in sorted(os.listdir(OBJ)) if f.endswith(".o")]
    libdir = os.path.join(ROOT, "fqzcomp5_b200")
    subprocess.run(["gcc", "-o", exe] + objs + ["-L" + libdir, "-lb200rans", "-Wl,-rpath," + libdir,
                                                "-lz", "-lm", "-pthread"], check=True)
    # the binary must get its codec from our library, not from a stray reference object
    nm = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    assert "rans_compress_4x16" in nm and "rans_uncompress_4x16" in nm
    return exe


@pytest.mark.parametrize("level", ["-1", "-3"])
def test_sample_fastq_is_byte_identical(fqz_gpu, golden, tmp_path, level):
    fq = os.path.join(ROOT, "tests", "golden", "sample.fastq")
    out, back = str(tmp_path / "s.fqz5"), str(tmp_path / "s.fq")
    subprocess.run([fqz_gpu, level, fq, out], check=True, capture_output=True)
    b = open(out, "rb").read()
    want = golden["fqzcomp5"][level]
    assert len(b) == want["len"] == 245
    assert hashlib.md5(b).hexdigest() == want["md5"]
    if level == "-1":
        assert want["md5"] == "8b5e07bf4c452ad206679f5e4bd7837a"      # SURVEY 8c known answer
    subprocess.run([fqz_gpu, "-d", out, back], check=True, capture_output=True)
    assert open(back, "rb").read() == open(fq, "rb").read()


def test_larger_fastq_through_stock_caller(fqz_gpu, tmp_path):
    """A block big enough to reach real rANS payloads (names, seq, qual; 4-lane),
    with several worker threads calling the library concurrently."""
    import numpy as np
    from fqzcomp5_b200 import synth
    n = 20000
    seq = synth.illumina_seq(n * 150).reshape(n, 150)
    qual = (synth.illumina_qual(n * 150) + 33).astype(np.uint8).reshape(n, 150)
    with open(tmp_path / "in.fq", "wb") as f:
        for i in range(n):
            f.write(b"@SIM.%d %d/1\n" % (i, i) + seq[i].tobytes() + b"\n+\n" + qual[i].tobytes() + b"\n")
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "fqzcomp5_ref")
    for level in ("-1", "-3"):
        out = str(tmp_path / ("o%s.fqz5" % level))
        subprocess.run([fqz_gpu, level, "-t", "4", str(tmp_path / "in.fq"), out], check=True, capture_output=True)
        if os.path.exists(ref_exe):
            rout = str(tmp_path / ("r%s.fqz5" % level))
            subprocess.run([ref_exe, level, "-t", "4", str(tmp_path / "in.fq"), rout], check=True,
                           capture_output=True)
            assert open(out, "rb").read() == open(rout, "rb").read(), "compressed file differs from the reference's"
        back = str(tmp_path / ("b%s.fq" % level))
        subprocess.run([fqz_gpu, "-d", "-t", "4", out, back], check=True, capture_output=True)
        assert open(back, "rb").read() == open(tmp_path / "in.fq", "rb").read()
