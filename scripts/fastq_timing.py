"""FASTQ split / join on the device next to the reference's load_seqs / output_fastq.

Prints one JSON line: device-resident GB/s of FASTQ text (CUDA events on the launching stream),
the HBM roofline fraction on algorithmic bytes (text read once + buffers written once, and the
reverse for join), the host-buffer C-ABI rate (copies inside the timed region), and the
reference's single-thread rate on a bounded sample of the same text (it parses a block on one
thread, fqzcomp5.c:2803-2815).

    python scripts/fastq_timing.py [records] [read_len]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                     # noqa: E402
from fqzcomp5_b200 import codec as bc, synth     # noqa: E402
from bench import load_peaks, cpu_fastq          # noqa: E402

def main(argv):
    nrec = int(argv[1]) if len(argv) > 1 else 3_000_000
    rl = int(argv[2]) if len(argv) > 2 else 150


    def make_text(nrec, rl):
        """@SIM.<9 digits> <9 digits>/1 \\n seq \\n + \\n qual \\n, fixed width so numpy can build it."""
        head = np.frombuffer(b"@SIM.000000000 000000000/1\n", np.uint8)
        w = head.size + rl + 1 + 2 + rl + 1
        rec = np.empty((nrec, w), np.uint8)
        rec[:, :head.size] = head
        idx = np.arange(nrec)
        for d in range(9):
            dig = ((idx // 10 ** (8 - d)) % 10 + 48).astype(np.uint8)
            rec[:, 5 + d] = dig
            rec[:, 15 + d] = dig
        o = head.size
        rec[:, o:o + rl] = synth.illumina_seq(nrec * rl).reshape(nrec, rl)
        rec[:, o + rl] = 10
        rec[:, o + rl + 1] = ord("+")
        rec[:, o + rl + 2] = 10
        rec[:, o + rl + 3:o + 2 * rl + 3] = (synth.illumina_qual(nrec * rl) + 33).reshape(nrec, rl)
        rec[:, w - 1] = 10
        return rec.reshape(-1)


    text = make_text(nrec, rl)
    n = int(text.size)
    dev = torch.device("cuda", 0)
    L = bc.lib()
    L.b200rans_set_device(0)
    host = bc.PinnedBuffer(n)
    host.array[:] = text
    d_text = torch.from_numpy(host.array).to(dev)
    mr = nrec + 16
    d_name = torch.empty(n // 4, dtype=torch.uint8, device=dev)
    d_seq = torch.empty(n // 2 + 64, dtype=torch.uint8, device=dev)
    d_qual = torch.empty(n // 2 + 64, dtype=torch.uint8, device=dev)
    d_len, d_flag, d_no, d_so = (torch.empty(mr, dtype=torch.int32, device=dev) for _ in range(4))
    sb = int(L.b200fq_split_scratch_bytes(n, mr))
    d_scr = torch.empty(sb + 256, dtype=torch.uint8, device=dev)
    scr = (d_scr.data_ptr() + 255) & ~255
    d_info = torch.zeros(16, dtype=torch.int32, device=dev)
    d_back = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream


    def split():
        rc = L.b200fq_split_dev(st, d_text.data_ptr(), n, d_name.data_ptr(), d_name.numel(), d_seq.data_ptr(),
                                d_qual.data_ptr(), d_seq.numel(), d_len.data_ptr(), d_flag.data_ptr(), d_no.data_ptr(),
                                d_so.data_ptr(), mr, scr, sb, d_info.data_ptr())
        assert rc == 0


    split()
    torch.cuda.synchronize()
    info = d_info.cpu().numpy().copy()
    assert info[0] == 0 and info[1] == nrec and info[6] == n, info
    name_len, seq_len = int(info[2]), int(info[3])
    sj = int(L.b200fq_join_scratch_bytes(name_len, nrec))
    d_scr2 = torch.empty(sj + 256, dtype=torch.uint8, device=dev)
    scr2 = (d_scr2.data_ptr() + 255) & ~255
    d_info2 = torch.zeros(16, dtype=torch.int32, device=dev)


    def join():
        rc = L.b200fq_join_dev(st, d_name.data_ptr(), name_len, d_seq.data_ptr(), d_qual.data_ptr(), d_len.data_ptr(),
                               nrec, 0, d_back.data_ptr(), d_back.numel(), scr2, sj, d_info2.data_ptr())
        assert rc == 0


    join()
    torch.cuda.synchronize()
    assert d_info2.cpu().numpy()[0] == 0 and int(d_info2.cpu().numpy()[7]) == n
    assert torch.equal(d_back[:n], d_text), "join(split(text)) != text"


    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        torch.cuda.synchronize()
        e[0].record()
        for i in range(reps):
            fn()
            e[i + 1].record()
        torch.cuda.synchronize()
        return float(np.mean([e[i].elapsed_time(e[i + 1]) for i in range(reps)]))


    ms_split, ms_join = timed(split), timed(join)
    peak, peak_src = load_peaks()
    out_bytes = name_len + 2 * seq_len + 8 * nrec
    alg = n + out_bytes

    # host-buffer C ABI
    name_h, seq_h, qual_h = bc.PinnedBuffer(n // 4), bc.PinnedBuffer(n // 2 + 64), bc.PinnedBuffer(n // 2 + 64)
    len_h, flag_h = np.empty(mr, np.uint32), np.empty(mr, np.uint32)
    hinfo = bc.FqInfo()
    tt = []
    for it in range(4):
        t0 = time.perf_counter()
        rc = L.b200fq_split(host.array.ctypes.data, n, name_h.array.ctypes.data, name_h.nbytes, seq_h.array.ctypes.data,
                            qual_h.array.ctypes.data, seq_h.nbytes, len_h.ctypes.data, flag_h.ctypes.data, mr,
                            C.addressof(hinfo))
        tt.append(time.perf_counter() - t0)
        assert rc == 0 and hinfo.status == 0 and hinfo.num_records == nrec
    e2e_split = min(tt[1:])
    back_h = bc.PinnedBuffer(n + 64)
    tt = []
    for it in range(4):
        t0 = time.perf_counter()
        rc = L.b200fq_join(name_h.array.ctypes.data, name_len, seq_h.array.ctypes.data, qual_h.array.ctypes.data, seq_len,
                           len_h.ctypes.data, nrec, 0, back_h.array.ctypes.data, back_h.nbytes, C.addressof(hinfo))
        tt.append(time.perf_counter() - t0)
        assert rc == 0 and hinfo.status == 0 and hinfo.text_len == n
    e2e_join = min(tt[1:])
    assert np.array_equal(back_h.array[:n], host.array)

    # the reference on one thread, bounded sample (bench.py owns every use of oracle/)
    w = n // nrec
    sample = np.ascontiguousarray(text[:w * min(nrec, 1_000_000)])
    t_split, t_join, kind = cpu_fastq(sample, sample.size // w)

    result = ({
        "workload": "synthetic Illumina FASTQ, %d records x %d bp, %d bytes of text" % (nrec, rl, n),
        "split": {"ms": ms_split, "gbs_text": n / ms_split / 1e6, "launches": 10,
                  "roofline": {"bound": "hbm", "achieved": alg / ms_split / 1e6, "peak": peak, "unit": "GB/s",
                               "frac": alg / ms_split / 1e6 / peak, "algorithmic_bytes": alg, "peak_source": peak_src}},
        "join": {"ms": ms_join, "gbs_text": n / ms_join / 1e6, "launches": 9,
                 "roofline": {"bound": "hbm", "achieved": alg / ms_join / 1e6, "peak": peak, "unit": "GB/s",
                              "frac": alg / ms_join / 1e6 / peak, "algorithmic_bytes": alg, "peak_source": peak_src}},
        "e2e": {"split_gbs": n / e2e_split / 1e9, "join_gbs": n / e2e_join / 1e9,
                "note": "host-buffer C ABI, pinned memory, H2D + kernels + D2H inside the timed region"},
        "cpu_baseline": {"kind": kind, "cores": 1, "split_gbs": sample.size / t_split / 1e9,
                         "join_gbs": sample.size / t_join / 1e9,
                         "sample": "first %d bytes of the same text, one thread (the reference parses a block on one "
                                   "thread); join writes to /dev/null" % sample.size}})
    return result


if __name__ == "__main__":
    print(json.dumps(main(sys.argv)))
