"""Method trial: b200rans_compress_methods_batch next to the reference's serial loop.

The reference's compress_with_methods (fqzcomp5.c:1979-2119) encodes one section buffer once per
candidate method and keeps the smallest; here all candidates of all inputs are one launch.  Prints one
JSON line per case: GB/s of section data trialled (input bytes / wall time, host buffers, copies inside
the timed region) on the GPU and on all host threads with the reference's SIMD build, after checking that
sizes, winners and winning streams agree on a sample.

    python scripts/trial_timing.py [bytes] [slice]
"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fqzcomp5_b200 import synth, codec as bc   # noqa: E402
from bench import cpu_codec                    # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 256_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 150 * 1747       # whole reads, ~256 KiB
cpu = cpu_codec()
cores = os.cpu_count() or 1

CASES = [
    # name, generator, methods (order values)
    ("qual -3 stock (4 lanes): RANS0,1,129,193", "illumina_qual", [0, 1, 129, 193]),
    ("qual -3 with X32:        RANS0,1,129,193 | 4", "illumina_qual", [4, 5, 133, 197]),
    ("seq stock (4 lanes):     RANS0,1,129,193", "illumina_seq", [0, 1, 129, 193]),
    ("seq with X32:            RANS0,1,129,193 | 4", "illumina_seq", [4, 5, 133, 197]),
]


def cpu_trial(buf, sl, methods, threads):
    base = buf.ctypes.data
    bound = max(cpu.bound(max(s for _, s in sl), m) for m in methods)
    idx, lock = [0], threading.Lock()
    res = [None] * len(sl)

    def worker():
        out = np.empty(bound + 16, np.uint8)
        while True:
            with lock:
                k = idx[0]
                idx[0] += 1
            if k >= len(sl):
                return
            o, s = sl[k]
            res[k] = [cpu.compress_into(base + o, s, out.ctypes.data, bound, m) for m in methods]
    th = [threading.Thread(target=worker) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return time.perf_counter() - t0, res


for name, gen, methods in CASES:
    host = bc.PinnedBuffer(total)
    host.array[:] = synth.GENERATORS[gen](total)
    sl = synth.slices(host.array, S)
    off = np.array([o for o, _ in sl], np.uint64)
    sz = np.array([s for _, s in sl], np.uint32)
    cap = int(sum(max(bc.rans_compress_bound_4x16(int(s), m) for m in methods) + 32 for s in sz[:1])) * len(sl) + 4096
    out = bc.PinnedBuffer(cap)
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        _, ooff, osz, best, csize = bc.compress_methods_batch(host.array, off, sz, methods, out=out.array)
        ts.append(time.perf_counter() - t0)
    tg = min(ts[1:])
    k = min(len(sl), max(cores * 4, 64))
    tc, res = cpu_trial(host.array, sl[:k], methods, cores)
    ok = all(list(csize[i]) == res[i] for i in range(k))
    ok = ok and all(int(best[i]) == int(np.argmin(res[i])) for i in range(k))
    sample = int(sz[:k].sum())
    print(json.dumps({
        "case": name, "methods": [hex(m) for m in methods], "inputs": len(sl), "slice_bytes": S,
        "bytes": total, "gpu_e2e_gbs": total / tg / 1e9, "gpu_candidate_gbs": total * len(methods) / tg / 1e9,
        "cpu_gbs": sample / tc / 1e9, "cpu_threads": cores, "cpu_sample_bytes": sample, "cpu_kind": cpu.kind,
        "sizes_and_winners_match_cpu": bool(ok),
        "winner_histogram": np.bincount(best, minlength=len(methods)).tolist(),
        "ratio": float(osz.sum()) / total}), flush=True)
    host.free()
    out.free()
