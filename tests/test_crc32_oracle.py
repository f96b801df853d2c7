"""The CRC-32 / block-framing oracle (oracle/crc32_oracle.c) against zlib itself (the library the
reference links for crc32()), the published check value, a block written by the unmodified
reference tool, and the committed known answer."""
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def crc_oracle():
    from oracle.pyoracle import Crc32Oracle
    return Crc32Oracle()


@pytest.fixture(scope="module")
def frame_golden():
    with open(os.path.join(ROOT, "tests", "golden", "block_frame.json")) as f:
        return json.load(f)


def test_crc32_matches_zlib_and_check_value(crc_oracle, frame_golden):
    g = frame_golden["crc32_check"]
    assert crc_oracle.crc32(g["input"].encode()) == int(g["crc"], 16) == zlib.crc32(g["input"].encode())
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 15, 16, 17, 511, 512, 513, 4096, 100000):
        b = rng.integers(0, 256, n).astype(np.uint8).tobytes()
        assert crc_oracle.crc32(b) == zlib.crc32(b)
        assert crc_oracle.crc32(b, 0x12345678) == zlib.crc32(b, 0x12345678)
    a, b = b"fqzcomp5 block ", b"framing"
    assert crc_oracle.crc32(b, crc_oracle.crc32(a)) == zlib.crc32(a + b)


def test_framing_matches_the_reference_tool(crc_oracle, frame_golden, tmp_path):
    """A block of a file written by the unmodified reference: re-framing its payload gives its bytes."""
    exe = os.path.join(ROOT, "oracle", "_ref", "fqzcomp5_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/fqzcomp5_ref not built")
    out = tmp_path / "s.fqz5"
    subprocess.run([exe, "-1", os.path.join(ROOT, "tests", "golden", "sample.fastq"), str(out)], check=True,
                   capture_output=True)
    d = out.read_bytes()
    g = frame_golden
    assert len(d) == g["file_bytes"]
    off = g["block_offset"]
    size, nrec, crc = struct.unpack_from("<III", d, off)
    assert (size, nrec, crc) == (g["block_size_field"], g["num_records"], int(g["crc_field"], 16))
    block = d[off:off + 4 + size]
    assert crc_oracle.frame_block(nrec, [block[12:40], block[40:]]) == block


def test_known_answer_without_the_reference(crc_oracle, frame_golden):
    """What the GPU box can check: the framing rule reproduces the recorded header fields for a
    payload whose CRC is known (the check string), and size = total - 4."""
    blk = crc_oracle.frame_block(7, [b"1234", b"56789"])
    size, nrec, crc = struct.unpack_from("<III", blk, 0)
    assert (size, nrec, crc) == (len(blk) - 4, 7, int(frame_golden["crc32_check"]["crc"], 16))
