"""CRC-32 on the device next to zlib's crc32 on one host thread (what the reference calls).

    python scripts/crc32_timing.py [bytes]
"""
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                  # noqa: E402
from fqzcomp5_b200 import codec as bc         # noqa: E402
from bench import load_peaks                  # noqa: E402

def main(argv):
    n = int(argv[1]) if len(argv) > 1 else 660_000_000      # a 1 GB quality block compresses to ~0.66 GB
    dev = torch.device("cuda", 0)
    L = bc.lib()
    L.b200rans_set_device(0)
    host = bc.PinnedBuffer(n)
    host.array[:] = np.random.default_rng(1).integers(0, 256, n, dtype=np.uint8)
    d = torch.from_numpy(host.array).to(dev)
    d_crc = torch.zeros(1, dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    for _ in range(3):
        assert L.b200fqz_crc32_dev(st, d.data_ptr() + 12, n - 12, 0, d_crc.data_ptr()) == 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(10):
        L.b200fqz_crc32_dev(st, d.data_ptr() + 12, n - 12, 0, d_crc.data_ptr())
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(10)]))
    got = int(d_crc.cpu().numpy().view(np.uint32)[0])
    t0 = time.perf_counter()
    want = zlib.crc32(host.array[12:])
    t1 = time.perf_counter()
    assert got == want, (hex(got), hex(want))
    peak, src = load_peaks()
    result = ({"workload": "crc32 over %d bytes at a 12-byte offset (an fqzcomp5 block after its CRC field)" % (n - 12),
                      "ms": ms, "gbs": (n - 12) / ms / 1e6, "launches": 3,
                      "roofline": {"bound": "hbm", "achieved": (n - 12) / ms / 1e6, "peak": peak, "unit": "GB/s",
                                   "frac": (n - 12) / ms / 1e6 / peak, "algorithmic_bytes": n - 12, "peak_source": src},
                      "cpu_baseline": {"kind": "reference dependency (zlib %s crc32, via Python's zlib module)" % zlib.ZLIB_VERSION,
                                       "cores": 1, "gbs": (n - 12) / (t1 - t0) / 1e9}})
    return result


if __name__ == "__main__":
    print(json.dumps(main(sys.argv)))
