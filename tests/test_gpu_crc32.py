"""CRC-32 and block framing on the device (fqzcomp5_b200/csrc/crc32.cu) against the oracle."""
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def crc_oracle():
    from oracle.pyoracle import Crc32Oracle
    return Crc32Oracle()


def test_crc32_host_buffers(gpu_codec, crc_oracle):
    assert gpu_codec.crc32(b"123456789") == 0xCBF43926
    rng = np.random.default_rng(2)
    for n in (0, 1, 15, 16, 17, 255, 511, 512, 513, 4095, 131071, 131072, 131073, 1 << 20, 3_000_001):
        b = rng.integers(0, 256, n).astype(np.uint8).tobytes()
        assert gpu_codec.crc32(b) == crc_oracle.crc32(b), n
        assert gpu_codec.crc32(b, 0xdeadbeef) == crc_oracle.crc32(b, 0xdeadbeef), n
    # running CRC over two pieces, as crc32(crc, buf, len) is used for appended data
    a, b = rng.integers(0, 256, 70001).astype(np.uint8).tobytes(), rng.integers(0, 256, 999).astype(np.uint8).tobytes()
    assert gpu_codec.crc32(b, gpu_codec.crc32(a)) == crc_oracle.crc32(a + b)


def test_crc32_device_any_alignment(gpu_codec, crc_oracle):
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    host = rng.integers(0, 256, 400000).astype(np.uint8)
    d = torch.from_numpy(host).to(dev)
    d_crc = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for off, n in [(0, 400000), (1, 1000), (12, 399988), (15, 17), (7, 131072), (13, 262144 + 5), (100, 0)]:
        assert gpu_codec.lib().b200fqz_crc32_dev(st, d.data_ptr() + off, n, 0, d_crc.data_ptr()) == 0
        torch.cuda.synchronize()
        assert int(d_crc.cpu().numpy().view(np.uint32)[0]) == crc_oracle.crc32(host[off:off + n].tobytes()), (off, n)


def test_assemble_block_from_device_streams(gpu_codec, crc_oracle, checker):
    """encode_block's framing with the seq / qual streams taken straight from the device-resident
    encoder's output and the small sections from host memory."""
    import torch
    import corpus
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    seq = np.frombuffer(corpus.make("illumina_seq", 150 * 2000, 1), np.uint8)
    qual = np.frombuffer(corpus.make("illumina_qual", 150 * 2000, 2), np.uint8)
    buf = np.concatenate([seq, qual])
    d_in = torch.from_numpy(buf).to(dev)
    in_off = np.array([0, seq.size], np.uint64); in_size = np.array([seq.size, qual.size], np.uint32)
    orders = np.array([0xC5, 5], np.int32)
    cap = gpu_codec.compress_bound_batch(in_size, orders)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(2, dtype=torch.int64, device=dev); d_sz = torch.zeros(2, dtype=torch.int32, device=dev)
    gpu_codec.compress_batch_dev(st, d_in.data_ptr(), in_off, in_size, orders, d_out.data_ptr(), cap,
                                 d_off.data_ptr(), d_sz.data_ptr())
    torch.cuda.synchronize()
    off, sz = d_off.cpu().numpy(), d_sz.cpu().numpy()
    cseq, cqual = checker.compress(seq.tobytes(), 0xC5), checker.compress(qual.tobytes(), 5)
    assert (int(sz[0]), int(sz[1])) == (len(cseq), len(cqual))
    names = np.frombuffer(b"\x01" + b"name-section-stand-in" * 11, np.uint8).copy()
    lens = np.frombuffer(bytes([2, 0x81, 0x16]), np.uint8).copy()          # fixed length 150 as encode_block writes it
    meta_s = np.frombuffer(struct.pack("<BII", 0, seq.size, len(cseq)), np.uint8).copy()
    meta_q = np.frombuffer(struct.pack("<BII", 0, qual.size, len(cqual)), np.uint8).copy()
    pieces = [(names.ctypes.data, names.size, 0), (lens.ctypes.data, lens.size, 0),
              (meta_s.ctypes.data, 9, 0), (d_out.data_ptr() + int(off[0]), int(sz[0]), 1),
              (meta_q.ctypes.data, 9, 0), (d_out.data_ptr() + int(off[1]), int(sz[1]), 1)]
    d_blk = torch.empty(12 + names.size + 3 + 18 + len(cseq) + len(cqual) + 64, dtype=torch.uint8, device=dev)
    n = gpu_codec.assemble_block_dev(st, 2000, pieces, d_blk.data_ptr(), d_blk.numel())
    torch.cuda.synchronize()
    got = d_blk[:n].cpu().numpy().tobytes()
    want = crc_oracle.frame_block(2000, [names.tobytes(), lens.tobytes(), meta_s.tobytes(), cseq, meta_q.tobytes(), cqual])
    assert got == want
