"""The C-ABI library: it loads, exports every symbol include/b200rans.h
declares, and fails loudly (never falls back) when there is no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    h = open(os.path.join(ROOT, "include", "b200rans.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b((?:rans|b200rans|b200fqz|b200fq)_\w+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from fqzcomp5_b200 import build, codec
    build.build()
    lib = ctypes.CDLL(codec.LIB_PATH)
    names = declared_functions()
    assert "rans_compress_to_4x16" in names and "b200rans_compress_batch" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(codec.EXPORTS) == names


def test_reference_symbols_are_the_ones_fqzcomp5_imports():
    """fqzcomp5.o / tokenise_name3.o import exactly these four."""
    names = declared_functions()
    for n in ("rans_compress_4x16", "rans_uncompress_4x16", "rans_compress_to_4x16",
              "rans_uncompress_to_4x16", "rans_compress_bound_4x16", "rans_set_cpu"):
        assert n in names


def test_bound_matches_reference_formula(oracle):
    from fqzcomp5_b200 import codec
    for n in (0, 1, 20, 21, 1000, 1001, 65536, 262144, 10 ** 6, 10 ** 8, 999999900, 2 ** 31 - 1):
        for order in (0, 1, 4, 5, 0x40, 0x80, 0xc5, 8, (150 << 8) | 9, (2 << 8) | 0xcd):
            assert codec.rans_compress_bound_4x16(n, order) == oracle.bound(n, order), (n, hex(order))


def test_no_cpu_fallback_without_gpu(capfd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device behaviour is checked on the CPU box")
    from fqzcomp5_b200 import codec
    assert codec.rans_compress_to_4x16(b"hello world, hello world", 0) is None
    assert codec.rans_uncompress_4x16(bytes([0x20, 3, 1, 2, 3])) is None
    err = capfd.readouterr().err
    assert "no usable CUDA device" in err or "no CPU path" in err
    import numpy as np
    buf = np.frombuffer(b"abcabcabc" * 100, np.uint8)
    with pytest.raises(codec.B200RansError):
        codec.compress_batch(buf, [0], [buf.size], [0])
    # the rows around the codec fail the same way: method trials, FASTQ split / join, CRC, framing
    with pytest.raises(codec.B200RansError):
        codec.compress_methods_batch(buf, [0], [buf.size], [0, 1])
    with pytest.raises(codec.B200RansError):
        codec.compress_trials(buf, [0], [buf.size], [[0, 1]])
    assert codec.compress_methods(b"hello world" * 10, [0, 1])[0] is None
    with pytest.raises(codec.B200RansError):
        codec.load_seqs(b"@r\nAC\n+\nII\n")
    with pytest.raises(codec.B200RansError):
        codec.output_fastq(b"r\0", b"AC", b"((", [2])
    with pytest.raises(codec.B200RansError):
        codec.crc32(b"123456789")


def test_new_entry_points_reject_bad_arguments_before_touching_a_device():
    """Argument validation that needs no GPU: NULL pointers and empty method lists are EINVAL / ENODEV,
    never a crash."""
    import ctypes as C
    from fqzcomp5_b200 import codec
    L = codec.lib()
    assert L.b200rans_compress_methods_batch(1, None, None, 0, None, None, 0, None, None, None, None) < 0
    assert L.b200rans_compress_trials(1, None, None, None, None, None, 0, None, None, None, None) < 0
    assert L.b200fq_split(None, 5, None, 0, None, None, 0, None, None, 0, None) < 0
    assert L.b200fq_join(None, 0, None, None, 0, None, 0, 0, None, 0, None) < 0
    assert L.b200fqz_crc32(0, None, 9, None) < 0
    n = C.c_uint(0)
    assert L.b200fqz_assemble_block_dev(None, 0, 1, None, None, 0, C.byref(n)) < 0
    assert L.b200fq_split_scratch_bytes(1 << 20, 1000) > 1000 * 16
    assert L.b200fq_join_scratch_bytes(1 << 20, 1000) > 1000 * 8


def test_product_never_imports_the_oracle():
    """oracle/ is a checker: nothing under fqzcomp5_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fqzcomp5_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_header_is_valid_c_and_the_example_links(tmp_path):
    """include/b200rans.h is a C header (the reference is C): the example in examples/ compiles as
    strict C99 and links against the library; without a GPU it fails loudly instead of falling back."""
    import subprocess
    import torch
    from fqzcomp5_b200 import build, codec
    build.build()
    exe = str(tmp_path / "block_pipeline")
    libdir = os.path.dirname(codec.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "block_pipeline.c"), "-L" + libdir, "-lb200rans",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "sample.fastq")], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "identical to" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode != 0 and "no CPU path" in r.stderr
