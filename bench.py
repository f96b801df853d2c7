#!/usr/bin/env python
"""bench.py -- rANS32x16 enc+dec throughput on B200 next to the reference's CPU codec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload illumina_qual_o0|ont_qual_o1|illumina_seq_c5|...]
                    [--bytes B] [--slice S]

A *step* is one pass of the hot path over one block of synthetic input: every
slice of the block is compressed (rans_compress_to_4x16 semantics) and the
streams are decompressed again.  `value` = uncompressed GB (1e9 B) per second of
the round trip (encode time + decode time), inputs resident in HBM; `e2e` is the
same through the host-buffer C ABI with H2D/D2H inside the timed region.

One JSON line on stdout (rank 0).  Multi-GPU: one process per GPU (torchrun),
each rank owns its own block (blocks are independent, no collective on the data
path); the barrier and the max-over-ranks time are the only communication.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fqzcomp5_b200 import synth  # noqa: E402

WORKLOADS = {
    # name: (generator, order, default bytes, description)
    "illumina_qual_o0": ("illumina_qual", 0x04, 999_999_900,
                         "synthetic Illumina 150bp qual stream, 1 GB block, rANS32x16 order-0 (configs[1])"),
    "illumina_qual_o1": ("illumina_qual", 0x05, 999_999_900,
                         "synthetic Illumina 150bp qual stream, 1 GB block, rANS32x16 order-1"),
    "illumina_seq_c5": ("illumina_seq", 0xC5, 999_999_900,
                        "synthetic Illumina seq stream, order-1 + PACK + RLE (configs[2])"),
    "ont_qual_o1": ("ont_qual", 0x05, 1_000_000_000,
                    "synthetic ONT long-read qual stream, rANS32x16 order-1 (configs[3])"),
}
METRIC = "rANS32x16 o0/o1 enc+dec GB/s (uncompressed) @1/2/4/8 B200 vs host CPU"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------ CPU arm
def cpu_codec():
    from oracle.pyoracle import Codec, available
    for kind in ("ref_simd", "ref", "oracle"):
        if available(kind) or kind == "oracle":
            try:
                return Codec(kind)
            except Exception:
                continue
    raise RuntimeError("no CPU checker available")


def cpu_roundtrip(codec, buf, slices, order, threads):
    """Time enc+dec of the given slices on `threads` host threads.  Returns (seconds_enc, seconds_dec, csize)."""
    import ctypes as C
    n = len(slices)
    bound = codec.bound(max(s for _, s in slices), order)
    outs = [np.empty(bound + 16, np.uint8) for _ in range(n)]
    back = [np.empty(s + 16, np.uint8) for _, s in slices]
    csz = [0] * n
    base = buf.ctypes.data

    def run(fn):
        idx = [0]
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    k = idx[0]
                    idx[0] += 1
                if k >= n:
                    return
                fn(k)
        th = [threading.Thread(target=worker) for _ in range(threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    def enc(k):
        o, s = slices[k]
        csz[k] = codec.compress_into(base + o, s, outs[k].ctypes.data, bound, order)

    def dec(k):
        o, s = slices[k]
        r = codec.uncompress_into(outs[k].ctypes.data, csz[k], back[k].ctypes.data, s)
        assert r == s
    te = run(enc)
    assert all(c > 0 for c in csz)
    td = run(dec)
    o, s = slices[0]
    assert bytes(back[0][:s]) == bytes(buf[o:o + s])
    return te, td, sum(csz)


def cpu_fastq(sample, nrec):
    """Reference load_seqs / output_fastq (oracle/_ref/libref_fqz.so, else the oracle port) on one thread
    over `sample` (uint8 array of FASTQ text): (split seconds, join seconds, kind).  Used by
    scripts/fastq_timing.py; raw C calls, no Python copies inside the timed regions."""
    import ctypes as C
    from oracle.pyoracle import FastqChecker, fastq_available, _libc, FqInfo
    if fastq_available("ref"):
        chk = FastqChecker("ref")
        last = C.c_int(0)
        t0 = time.perf_counter()
        fq = chk.lib.load_seqs(sample.ctypes.data, int(sample.size), C.byref(last))
        t1 = time.perf_counter()
        assert fq and fq.contents.num_records == nrec
        qb = np.ctypeslib.as_array(C.cast(fq.contents.qual_buf, C.POINTER(C.c_ubyte)), (fq.contents.qual_len,))
        fp = _libc.fopen(b"/dev/null", b"wb")
        t2 = time.perf_counter()
        qb += 33                                                   # fqzcomp5.c:2532-2533
        chk.lib.output_fastq(fp, fq, 0)
        t3 = time.perf_counter()
        _libc.fclose(fp)
        chk.lib.fastq_free(fq)
        return t1 - t0, t3 - t2, "reference"
    chk = FastqChecker("oracle")
    nm, sq, ql = (np.empty(sample.size, np.uint8) for _ in range(3))
    ln, fl = np.empty(sample.size // 4, np.uint32), np.empty(sample.size // 4, np.uint32)
    oi = FqInfo()
    t0 = time.perf_counter()
    chk.lib.fqo_split(sample.ctypes.data, int(sample.size), nm.ctypes.data, sq.ctypes.data, ql.ctypes.data,
                      ln.ctypes.data, fl.ctypes.data, C.byref(oi))
    t1 = time.perf_counter()
    out = np.empty(sample.size + 64, np.uint8)
    chk.lib.fqo_join(nm.ctypes.data, sq.ctypes.data, ql.ctypes.data, ln.ctypes.data, oi.num_records, 0,
                     out.ctypes.data)
    return t1 - t0, time.perf_counter() - t1, "port"


def kind_name(codec):
    return "reference" if codec.kind.startswith("ref") else "port"


def run_reference(args, gen, order, total, S):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    codec = cpu_codec()
    cores = os.cpu_count() or 1
    # bounded sample of the same workload: ~64 MiB per host thread, capped
    sample = int(min(total, max(S, min(cores, 64) * (32 << 20))))
    buf = synth.GENERATORS[gen](sample)
    sl = synth.slices(buf, S)
    for _ in range(args.warmup):
        cpu_roundtrip(codec, buf, sl[:max(cores, 1)], order, cores)
    tt = []
    csize = 0
    for _ in range(args.steps):
        te, td, csize = cpu_roundtrip(codec, buf, sl, order, cores)
        tt.append((te, td))
    te = float(np.mean([a for a, _ in tt]))
    td = float(np.mean([b for _, b in tt]))
    gbs = sample / (te + td) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": (te + td) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": WORKLOADS[args.workload][3],
                   "order": hex(order), "block_bytes": total, "slice_bytes": S, "streams": len(sl),
                   "sample_bytes": sample,
                   "note": "CPU arm: each step codes a bounded sample of the block on all host threads"},
        "enc_gbs": sample / te / 1e9, "dec_gbs": sample / td / 1e9, "ratio": csize / sample,
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind_name(codec),
                         "sample": "%d slices of %d B (%.0f MB) of the workload, %s build, %d threads"
                                   % (len(sl), S, sample / 1e6, codec.kind, cores)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------ GPU arm
def run_b200(args, gen, order, total, S):
    import torch
    import torch.distributed as dist
    from fqzcomp5_b200 import codec as bc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    bc.lib().b200rans_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- one block per rank (weak scaling; blocks are independent)
    host = bc.PinnedBuffer(total)
    host.array[:] = synth.GENERATORS[gen](total, seed=synth_seed(gen) + 7919 * rank)
    sl = synth.slices(host.array, S)
    n = len(sl)
    in_off = np.array([o for o, _ in sl], np.uint64)
    in_size = np.array([s for _, s in sl], np.uint32)
    orders = np.full(n, order, np.int32)
    d_in = torch.empty(total, dtype=torch.uint8, device=dev)
    d_in.copy_(torch.from_numpy(host.array), non_blocking=False)
    cap = bc.compress_bound_batch(in_size, orders)
    d_comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_coff = torch.zeros(n, dtype=torch.int64, device=dev)
    d_csz = torch.zeros(n, dtype=torch.int32, device=dev)
    d_back = torch.empty(total, dtype=torch.uint8, device=dev)
    d_osz = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    # a dedicated (non-default) stream: the library launches on it and the timing events are recorded on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    bc.set_profiling(True)

    def enc():
        bc.compress_batch_dev(stream, d_in.data_ptr(), in_off, in_size, orders, d_comp.data_ptr(), cap,
                              d_coff.data_ptr(), d_csz.data_ptr())

    def dec(coff, csz, flags):
        bc.uncompress_batch_dev(stream, d_comp.data_ptr(), coff, csz, d_back.data_ptr(), in_off, in_size,
                                d_osz.data_ptr(), d_st.data_ptr(), flags=flags)

    # first pass: sizes, flags, correctness of the round trip
    enc()
    torch.cuda.synchronize()
    coff = d_coff.cpu().numpy().astype(np.uint64)
    csz = d_csz.cpu().numpy().astype(np.uint32)
    assert (csz > 0).all(), "a stream failed to compress"
    flags = d_comp[torch.from_numpy(coff.astype(np.int64)).to(dev)].cpu().numpy()
    dec(coff, csz, flags)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0, "a stream failed to decompress"
    assert torch.equal(d_back, d_in), "round trip mismatch"
    csize = int(csz.sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        enc()
        dec(coff, csz, flags)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # keep the GPU under the same load for a moment so the clock samples cover the timed region
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        enc()
        dec(coff, csz, flags)
        torch.cuda.synchronize()
    launches0 = bc.launch_count()
    t_enc, t_dec, k_enc, k_dec = [], [], [], []
    e0, e1, e2 = ev(), ev(), ev()
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        e0.record()
        enc()
        e1.record()
        dec(coff, csz, flags)
        e2.record()
        e2.synchronize()
        t_enc.append(e0.elapsed_time(e1))
        t_dec.append(e1.elapsed_time(e2))
        k_enc.append(bc.last_kernel_ms(0))
        k_dec.append(bc.last_kernel_ms(1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = bc.launch_count() - launches0
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        enc()
        dec(coff, csz, flags)
        torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_enc, ms_dec = float(np.mean(t_enc)), float(np.mean(t_dec))
    ms_step = ms_enc + ms_dec

    # ---- e2e: the host-buffer C ABI, pinned host memory, copies inside the timed region
    out_host = bc.PinnedBuffer(cap + 4096)
    back_host = bc.PinnedBuffer(total)
    e2e_t = []
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            barrier()
        t0 = time.perf_counter()
        _, ooff, osz = bc.compress_batch(host.array, in_off, in_size, orders, out=out_host.array)
        t1 = time.perf_counter()
        bc.uncompress_batch(out_host.array, ooff, osz, back_host.array, in_off, in_size)
        t2 = time.perf_counter()
        if it >= args.warmup:
            e2e_t.append((t1 - t0, t2 - t1))
    assert np.array_equal(back_host.array, host.array), "e2e round trip mismatch"
    e2e_enc = float(np.mean([a for a, _ in e2e_t]))
    e2e_dec = float(np.mean([b for _, b in e2e_t]))

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_step, ms_enc, ms_dec, (e2e_enc + e2e_dec) * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_enc, ms_dec, e2e_ms = [float(x) for x in t.tolist()]
    else:
        e2e_ms = (e2e_enc + e2e_dec) * 1e3

    if rank == 0:
        peak, peak_src = load_peaks()
        U, Cc = total, csize
        kd, ke = float(np.mean(k_dec)), float(np.mean(k_enc))

        traffic = {}
        try:        # DRAM bytes per launch from the committed ncu --set full capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if args.workload in tj and total == WORKLOADS[args.workload][2] and S == (256 << 10):
                traffic = tj[args.workload]
        except Exception:
            pass

        def roof(ms, name):
            a = (U + Cc) / (ms * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s",
                    "frac": a / peak, "traffic": traffic.get(name), "traffic_source": traffic.get("source"),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": U + Cc, "kernel_ms": ms}
        r_enc, r_dec = roof(ke, "enc_kernel"), roof(kd, "dec_kernel")
        dominant = r_enc if ke >= kd else r_dec
        line = {
            "metric": METRIC, "value": world * U / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": WORKLOADS[args.workload][3],
                       "order": hex(order), "block_bytes": U, "slice_bytes": S, "streams": n,
                       "blocks": world, "l2": "inputs (%.2f GB per step) exceed the 126 MB L2" % (U / 1e9),
                       "ratio": Cc / U},
            "enc_gbs": world * U / (ms_enc * 1e-3) / 1e9, "dec_gbs": world * U / (ms_dec * 1e-3) / 1e9,
            "e2e": {"value": world * U / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": U + Cc, "d2h_bytes_per_step": U + Cc,
                    "enc_gbs": U / e2e_enc / 1e9, "dec_gbs": U / e2e_dec / 1e9},
            "gpu_launches": int(launches),
            "roofline": dominant, "roofline_enc": r_enc, "roofline_dec": r_dec,
            "clocks": sampler.summary(), "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu:
            codec = cpu_codec()
            cores = os.cpu_count() or 1
            sample = int(min(total, max(S, min(cores, 64) * (16 << 20))))
            k = max(1, sample // S)
            te, td, cs = cpu_roundtrip(codec, host.array, sl[:k], order, cores)
            sb = sum(s for _, s in sl[:k])
            line["cpu_baseline"] = {
                "value": sb / (te + td) / 1e9, "unit": "GB/s", "cores": cores, "kind": kind_name(codec),
                "sample": "first %d slices of %d B (%.0f MB) of the same block, %s build, %d threads"
                          % (k, S, sb / 1e6, codec.kind, cores),
                "enc_gbs": sb / te / 1e9, "dec_gbs": sb / td / 1e9}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------ rows around the codec
EXTRA = {
    "fastq_split": "FASTQ block split (load_seqs) GB/s of text on one B200",
    "fastq_join": "FASTQ block join (output_fastq) GB/s of text on one B200",
    "crc32": "CRC-32 of a compressed block GB/s on one B200",
}


def run_extra(args):
    """The steps either side of the codec, measured by scripts/fastq_timing.py and
    scripts/crc32_timing.py (CUDA events on the launching stream, inputs resident in HBM, several GB of
    distinct data per timed loop so nothing is served from L2), reported in this file's JSON shape."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    if args.workload == "crc32":
        import crc32_timing
        r = crc32_timing.main(["crc32_timing"] + ([str(args.bytes)] if args.bytes else []))
        value, ms, roof = r["gbs"], r["ms"], r["roofline"]
        e2e = None
        cpu = {"value": r["cpu_baseline"]["gbs"], "unit": "GB/s", "cores": 1, "kind": "reference",
               "sample": r["cpu_baseline"]["kind"] + ", the whole buffer"}
        launches = r["launches"]
    else:
        import fastq_timing
        r = fastq_timing.main(["fastq_timing"] + ([str(args.bytes // 331)] if args.bytes else []))
        k = "split" if args.workload == "fastq_split" else "join"
        value, ms, roof = r[k]["gbs_text"], r[k]["ms"], r[k]["roofline"]
        e2e = {"value": r["e2e"][k + "_gbs"], "unit": "GB/s", "note": r["e2e"]["note"]}
        cpu = {"value": r["cpu_baseline"][k + "_gbs"], "unit": "GB/s", "cores": 1, "kind": r["cpu_baseline"]["kind"],
               "sample": r["cpu_baseline"]["sample"]}
        launches = r[k]["launches"]
    roof = dict(roof, traffic=None)
    emit({"metric": EXTRA[args.workload], "value": value, "unit": "GB/s", "n_gpus": 1, "steps": 10, "warmup": 3,
          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
          "data": "synthetic", "config": {"workload": args.workload, "description": r["workload"],
                                          "l2": "inputs exceed the 126 MB L2"},
          "e2e": e2e, "gpu_launches": launches * 10, "roofline": roof, "cpu_baseline": cpu})


def synth_seed(gen):
    return {"illumina_qual": 2, "binned_qual": 22, "illumina_seq": 3, "ont_qual": 4}[gen]


class QuietStdout:
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr, so that
    stdout carries exactly one JSON line; emit() writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


OUT = None


def emit(obj):
    line = json.dumps(obj)
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line)


def main():
    global OUT
    OUT = QuietStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="illumina_qual_o0", choices=sorted(WORKLOADS) + sorted(EXTRA))
    ap.add_argument("--bytes", type=int, default=0, help="block size per GPU (default: the config's 1 GB)")
    ap.add_argument("--slice", type=int, default=256 << 10, help="bytes per rans_compress_to_4x16 call")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.workload in EXTRA:
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                emit({"impl": "reference", "unavailable": "the CPU leg of %s is reported inside the b200 arm's "
                      "cpu_baseline (one thread: the reference runs this step on one thread)" % args.workload})
            return
        return run_extra(args)
    gen, order, total, _ = WORKLOADS[args.workload]
    if args.bytes:
        total = args.bytes
    if args.impl == "reference":
        run_reference(args, gen, order, total, args.slice)
    else:
        run_b200(args, gen, order, total, args.slice)


if __name__ == "__main__":
    main()
