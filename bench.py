#!/usr/bin/env python
"""bench.py -- rANS32x16 enc+dec throughput on B200 next to the reference's CPU codec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload all|illumina_qual_o0|illumina_qual_o1|illumina_seq_c5|ont_qual_o1|
                                fastq_blocks_m3|fastq_split|fastq_join|crc32]
                    [--bytes B] [--slice S] [--blocks-per-gpu B]

A *step* is one pass of the hot path over one block of synthetic input: every slice of the block is
compressed (rans_compress_to_4x16 semantics) and the streams are decompressed again.  `value` =
uncompressed GB (1e9 B) per second of the round trip (encode time + decode time), inputs resident in
HBM; `e2e` is the same through the host-buffer C ABI with H2D/D2H inside the timed region.

The default (`--workload all`) line is the headline workload (configs[1], illumina_qual_o0) with the other
configs under `per_config`: configs[2] illumina_seq_c5, configs[3] ont_qual_o1, illumina_qual_o1 (N = 1
only) and configs[4] fastq_blocks_m3 -- FASTQ blocks through b200fqz_encode_blocks_multi /
decode_blocks_multi with fqzcomp5 -3's method sets, dealt over the N GPUs by the library's persistent
workers in ONE process (rank 0), stock 4-lane lists and with RANS_ORDER_X32.

One JSON line on stdout (rank 0).  Multi-GPU: one process per GPU (torchrun), each rank owns its own
block (blocks are independent, no collective on the data path); a gloo barrier and the max-over-ranks
time are the only communication.  NCCL is not used.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fqzcomp5_b200 import partition, synth  # noqa: E402

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
HEADLINE = "illumina_qual_o0"
PER_CONFIG = ["illumina_seq_c5", "ont_qual_o1", "illumina_qual_o1"]
BLOCKS = "fastq_blocks_m3"
METRIC = "rANS32x16 o0/o1 enc+dec GB/s (uncompressed) @1/2/4/8 B200 vs host CPU"
# fqzcomp5 -3 (fqzcomp5.c:4893-4900): the rANS members of its seq / qual method sets; names: the codec half of
# TLZP3 (order 5).  LZP and tok3 are host stages outside the path.
M3_SEQ, M3_QUAL, M3_NAMES = (0, 1, 129, 193), (0, 1, 129, 193, -1), (5,)
BLOCK_RECORDS, READ_LEN = 3_021_148, 150           # 3 021 148 x 331 B = 999 999 988 B of FASTQ text


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


def pcie_ceiling(world, ratio):
    """Round-trip ceiling of the host-buffer path at `world` GPUs from the committed probe of this pool's boxes
    (scripts/pcie_probe.py -> profiles/pcie_probe.jsonl: all GPUs copying at once).  Encode moves U in and C out
    at the same time, decode C in and U out; each direction is limited by the aggregate rate the probe reached in
    that direction, so per unit of U
        t >= max(1 / H2D, ratio / D2H)     (encode; decode with H2D and D2H swapped)
    and the round trip by the sum of the two.  (With both directions busy the probe reaches less per direction;
    that figure is carried along, it is not a bound for unequal traffic.)  None when no probe is committed for
    this GPU count."""
    try:
        rows = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "pcie_probe.jsonl")) if l.strip().startswith("{")]
        r = [x for x in rows if x["gpus"] == world]
        if not r:
            return None
        h, d, b = r[0]["h2d_gbs_total"], r[0]["d2h_gbs_total"], r[0]["both_gbs_each_direction_total"]
        t_enc = max(1.0 / h, ratio / d)
        t_dec = max(ratio / h, 1.0 / d)
        lim = lambda t, a, c: "H2D" if t == a else "D2H"
        return {"h2d_gbs_total": h, "d2h_gbs_total": d, "both_gbs_each_direction_total": b,
                "enc_ceiling_gbs": 1.0 / t_enc, "dec_ceiling_gbs": 1.0 / t_dec,
                "round_trip_ceiling_gbs": 1.0 / (t_enc + t_dec),
                "limit": {"enc": lim(t_enc, 1.0 / h, ratio / d), "dec": lim(t_dec, ratio / h, 1.0 / d)},
                "source": "profiles/pcie_probe.jsonl (scripts/pcie_probe.py on this pool's 8-GPU box)"}
    except Exception:
        return None


def codec_config(workload, order, U, S, n, world, ratio):
    """`config` of a codec workload: the same dict in the b200 and the reference arm."""
    return {"workload": workload, "description": WORKLOADS[workload][3], "order": hex(order), "block_bytes": U,
            "slice_bytes": S, "streams": n, "blocks": world,
            "l2": "inputs (%.2f GB per step) exceed the 126 MB L2" % (U / 1e9), "ratio": ratio,
            "output": "one rans_compress_bound_4x16-sized buffer per call (in-slot), as the reference's callers "
                      "provide; decode flags (first byte of each stream) are host-known",
            "issue": "the timed steps are issued back to back; one synchronize closes the timed region"}


def synth_seed(gen):
    return {"illumina_qual": 2, "binned_qual": 22, "illumina_seq": 3, "ont_qual": 4}[gen]


# ------------------------------------------------------------------ CPU arm
def cpu_codec(kind=None):
    from oracle.pyoracle import Codec, available
    for k in ([kind] if kind else ["ref_simd", "ref", "oracle"]):
        if available(k) or k == "oracle":
            try:
                return Codec(k)
            except Exception:
                continue
    raise RuntimeError("no CPU checker available")


def run_threads(n, fn, threads):
    """fn(k) for k in range(n) on `threads` host threads (ctypes calls release the GIL).  Returns seconds."""
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


def cpu_roundtrip(codec, buf, slices, order, threads, keep=False):
    """Time enc+dec of the given slices on `threads` host threads.
    Returns (seconds_enc, seconds_dec, csize[, list of stream bytes])."""
    n = len(slices)
    bound = codec.bound(max(s for _, s in slices), order)
    outs = [np.empty(bound + 16, np.uint8) for _ in range(n)]
    back = [np.empty(s + 16, np.uint8) for _, s in slices]
    csz = [0] * n
    base = buf.ctypes.data

    def enc(k):
        o, s = slices[k]
        csz[k] = codec.compress_into(base + o, s, outs[k].ctypes.data, bound, order)

    def dec(k):
        o, s = slices[k]
        r = codec.uncompress_into(outs[k].ctypes.data, csz[k], back[k].ctypes.data, s)
        assert r == s
    te = run_threads(n, enc, threads)
    assert all(c > 0 for c in csz)
    td = run_threads(n, dec, threads)
    o, s = slices[0]
    assert bytes(back[0][:s]) == bytes(buf[o:o + s])
    if keep:
        return te, td, sum(csz), [outs[k][:csz[k]] for k in range(n)]
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


def run_reference(args, name):
    """--impl reference: the reference's own CPU codec (AVX2/AVX-512 build) on all host threads over the
    WHOLE block of the workload, same slices, same config keys as the b200 arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gen, order, total, _ = WORKLOADS[name]
    if args.bytes:
        total = args.bytes
    S = args.slice
    codec = cpu_codec()
    cores = os.cpu_count() or 1
    synth.PROCESSES = min(cores, 32)
    buf = synth.GENERATORS[gen](total, seed=synth_seed(gen))
    sl = synth.slices(buf, S)
    for _ in range(args.warmup):
        cpu_roundtrip(codec, buf, sl[:max(4 * cores, 1)], order, cores)
    tt = []
    csize = 0
    for _ in range(args.steps):
        te, td, csize = cpu_roundtrip(codec, buf, sl, order, cores)
        tt.append((te, td))
    te = float(np.mean([a for a, _ in tt]))
    td = float(np.mean([b for _, b in tt]))
    gbs = total / (te + td) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": (te + td) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32",
        "data": "synthetic",
        "config": codec_config(name, order, total, S, len(sl), max(args.gpus, 1), csize / total),
        "enc_gbs": total / te / 1e9, "dec_gbs": total / td / 1e9,
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind_name(codec),
                         "sample": "the whole block: %d slices of %d B (%.0f MB), %s build, %d threads (one rank "
                                   "codes one block whatever --gpus says: the host is shared)"
                                   % (len(sl), S, total / 1e6, codec.kind, cores)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------ GPU arm: one codec workload
def measure_codec(args, name, data, env, with_cpu):
    """One codec workload on this rank's GPU.  data: uint8 numpy array (the block).  Returns the result dict
    (times are this rank's; the caller reduces over ranks)."""
    torch, bc, dev, stream = env["torch"], env["bc"], env["dev"], env["stream"]
    gen, order, _, _ = WORKLOADS[name]
    S = args.slice
    total = int(data.size)
    host = bc.PinnedBuffer(total)
    host.array[:] = data
    sl = synth.slices(host.array, S)
    n = len(sl)
    in_off = np.array([o for o, _ in sl], np.uint64)
    in_size = np.array([s for _, s in sl], np.uint32)
    orders = np.full(n, order, np.int32)
    d_in = torch.empty(total, dtype=torch.uint8, device=dev)
    d_in.copy_(torch.from_numpy(host.array), non_blocking=False)
    cap = bc.compress_slots_bound(in_size, orders)
    d_comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_coff = torch.zeros(n, dtype=torch.int64, device=dev)
    d_csz = torch.zeros(n, dtype=torch.int32, device=dev)
    d_back = torch.empty(total, dtype=torch.uint8, device=dev)
    d_osz = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    bc.set_profiling(True)

    def enc():      # one bound-sized output buffer per call, as the reference's callers provide (in-slot)
        bc.compress_batch_dev2(stream, d_in.data_ptr(), in_off, in_size, orders, d_comp.data_ptr(), cap,
                               d_coff.data_ptr(), d_csz.data_ptr(), flags=bc.OUT_IN_SLOT)

    def enc_packed():   # streams packed back to back into one arena (scan + gather after the coder)
        bc.compress_batch_dev(stream, d_in.data_ptr(), in_off, in_size, orders, d_comp.data_ptr(), cap,
                              d_coff.data_ptr(), d_csz.data_ptr())

    def dec(coff, csz, flags):
        bc.uncompress_batch_dev(stream, d_comp.data_ptr(), coff, csz, d_back.data_ptr(), in_off, in_size,
                                d_osz.data_ptr(), d_st.data_ptr(), flags=flags)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # packed form first (a few passes, reported beside the headline), then the in-slot form stays in d_comp
    tp = []
    for it in range(1 + min(args.steps, 3)):
        a, b = ev(), ev()
        a.record()
        enc_packed()
        b.record()
        b.synchronize()
        if it:
            tp.append(a.elapsed_time(b))
    packed_csz = d_csz.cpu().numpy().astype(np.uint32)
    # first pass: sizes, flags, correctness of the round trip
    enc()
    torch.cuda.synchronize()
    coff = d_coff.cpu().numpy().astype(np.uint64)
    csz = d_csz.cpu().numpy().astype(np.uint32)
    assert (csz > 0).all(), "a stream failed to compress"
    assert np.array_equal(csz, packed_csz), "in-slot and packed outputs differ in size"
    flags = d_comp[torch.from_numpy(coff.astype(np.int64)).to(dev)].cpu().numpy()
    flag_hist = {hex(int(f)): int(c) for f, c in zip(*np.unique(flags, return_counts=True))}
    if name == "illumina_seq_c5":
        # the emitted flag byte keeps PACK (0x80) -- every stream -- and RLE (0x40).  At 256 KiB
        # per call the reference itself rejects RLE for about a third of the slices (its .99 test,
        # rANS_static4x16pr.c:1485: too few poly-G tails in that slice); those streams are 0x85 on the CPU too
        # (the byte comparison below covers both kinds).
        assert ((flags & 0x85) == 0x85).all(), "config 3: a stream dropped PACK, X32 or order 1"
        assert int(((flags & 0xC0) == 0xC0).sum()) * 2 > flags.size, "config 3: RLE kept by fewer than half the streams"
    dec(coff, csz, flags)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0, "a stream failed to decompress"
    assert torch.equal(d_back, d_in), "round trip mismatch"
    csize = int(csz.sum())

    for _ in range(args.warmup):
        enc()
        dec(coff, csz, flags)
    env["barrier"]()
    sampler = ClockSampler(env["local"])
    sampler.start()
    # keep the GPU under the same load for a moment so the clock samples cover the timed region
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        enc()
        dec(coff, csz, flags)
        torch.cuda.synchronize()
    launches0 = bc.launch_count()
    t_enc, t_dec, k_enc, k_dec = [], [], [], []
    # The K timed steps are queued back to back, as a caller with many blocks would issue them: the calls are
    # asynchronous, so the host's planning of a call (job records, their upload) runs while the device works on the
    # previous one.  One event after every call, read after the synchronize that closes the timed region.
    evs = [ev() for _ in range(2 * args.steps + 1)]
    env["barrier"]()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    evs[0].record()
    for k in range(args.steps):
        enc()
        evs[2 * k + 1].record()
        dec(coff, csz, flags)
        evs[2 * k + 2].record()
    torch.cuda.synchronize()
    env["barrier"]()
    wall = time.perf_counter() - wall0
    for k in range(args.steps):
        t_enc.append(evs[2 * k].elapsed_time(evs[2 * k + 1]))
        t_dec.append(evs[2 * k + 1].elapsed_time(evs[2 * k + 2]))
    launches = bc.launch_count() - launches0
    # the coder kernels alone (events the library records around them; one pair per context, so these extra steps
    # are read one at a time) -- the GPU stays under the same load for the clock sampler
    for _ in range(args.steps):
        enc()
        dec(coff, csz, flags)
        torch.cuda.synchronize()
        k_enc.append(bc.last_kernel_ms(0))
        k_dec.append(bc.last_kernel_ms(1))
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        enc()
        dec(coff, csz, flags)
        torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # the steps issued back to back left the same bytes as the checked first pass
    assert int(d_st.abs().sum()) == 0 and torch.equal(d_back, d_in), "round trip mismatch after the timed steps"
    assert np.array_equal(d_csz.cpu().numpy().astype(np.uint32), csz), "stream sizes changed during the timed steps"
    ms_enc, ms_dec = float(np.mean(t_enc)), float(np.mean(t_dec))
    # streams of the first slices, as the device left them (for the byte comparison with the CPU below)
    kpar = min(n, 64)
    dev_streams = [d_comp[int(coff[k]):int(coff[k]) + int(csz[k])].cpu().numpy() for k in range(kpar)]
    del d_comp, d_back

    # ---- e2e: the host-buffer C ABI, pinned host memory, copies inside the timed region
    out_host = bc.PinnedBuffer(bc.compress_bound_batch(in_size, orders) + 4096)
    back_host = bc.PinnedBuffer(total)
    e2e_t = []
    ooff = osz = None
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            env["barrier"]()
        t0 = time.perf_counter()
        _, ooff, osz = bc.compress_batch(host.array, in_off, in_size, orders, out=out_host.array)
        t1 = time.perf_counter()
        bc.uncompress_batch(out_host.array, ooff, osz, back_host.array, in_off, in_size)
        t2 = time.perf_counter()
        if it >= args.warmup:
            e2e_t.append((t1 - t0, t2 - t1))
    assert np.array_equal(back_host.array, host.array), "e2e round trip mismatch"
    assert np.array_equal(osz, csz), "host-buffer and device-resident sizes differ"
    e2e_enc = float(np.mean([a for a, _ in e2e_t]))
    e2e_dec = float(np.mean([b for _, b in e2e_t]))

    U, Cc = total, csize
    kd, ke = float(np.mean(k_dec)), float(np.mean(k_enc))
    peak, peak_src = load_peaks()
    traffic = {}
    try:        # DRAM bytes per launch from the committed ncu --set full captures of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if name in tj and total == WORKLOADS[name][2] and S == (256 << 10):
            traffic = tj[name]
    except Exception:
        pass

    def roof(ms, kname):
        a = (U + Cc) / (ms * 1e-3) / 1e9
        return {"kernel": kname, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s",
                "frac": a / peak, "traffic": traffic.get(kname), "traffic_source": traffic.get("source"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": U + Cc, "kernel_ms": ms}
    r_enc, r_dec = roof(ke, "enc_kernel"), roof(kd, "dec_kernel")
    # step level: everything the encode / decode call launches (histogram, coder, results), on U + C
    step = {"enc_step_ms": ms_enc, "dec_step_ms": ms_dec,
            "enc_step_frac": (U + Cc) / (ms_enc * 1e-3) / 1e9 / peak,
            "dec_step_frac": (U + Cc) / (ms_dec * 1e-3) / 1e9 / peak,
            "enc_step_traffic": traffic.get("enc_step"), "dec_step_traffic": traffic.get("dec_step")}
    res = {
        "ms_enc": ms_enc, "ms_dec": ms_dec, "e2e_enc_s": e2e_enc, "e2e_dec_s": e2e_dec,
        "U": U, "C": Cc, "n": n, "order": order, "S": S, "flags": flag_hist,
        "packed_enc_ms": float(np.mean(tp)) if tp else None,
        "gpu_launches": int(launches), "roofline": r_enc if ke >= kd else r_dec,
        "roofline_enc": r_enc, "roofline_dec": r_dec, "roofline_step": step,
        "clocks": sampler.summary(), "wall_s_timed_region": wall,
    }
    if with_cpu:
        cores = os.cpu_count() or 1
        codec = cpu_codec()
        sample = int(min(total, max(S, min(cores, 64) * (16 << 20))))
        k = max(1, sample // S)
        te, td, cs, couts = cpu_roundtrip(codec, host.array, sl[:k], order, cores, keep=True)
        sb = sum(s for _, s in sl[:k])
        # parity on the timed data: the streams of the slices the CPU leg just coded, byte for byte, against
        # what the host-buffer call returned and what the device-resident call left in its slots
        bad = 0
        for j in range(k):
            g = out_host.array[int(ooff[j]):int(ooff[j]) + int(osz[j])]
            if g.size != couts[j].size or not np.array_equal(g, couts[j]):
                bad += 1
            if j < kpar and not np.array_equal(dev_streams[j], couts[j]):
                bad += 1
        assert bad == 0, "%d of the first %d streams differ from the CPU codec's" % (bad, k)
        res["parity_checked_slices"] = k
        res["cpu_baseline"] = {
            "value": sb / (te + td) / 1e9, "unit": "GB/s", "cores": cores, "kind": kind_name(codec),
            "sample": "first %d slices of %d B (%.0f MB) of the same block, %s build, %d threads"
                      % (k, S, sb / 1e6, codec.kind, cores),
            "enc_gbs": sb / te / 1e9, "dec_gbs": sb / td / 1e9}
        try:        # the as-shipped build (cpuid macros undefined: scalar 32x16 dispatch)
            sc = cpu_codec("ref")
            k2 = max(1, k // 4)
            te2, td2, _ = cpu_roundtrip(sc, host.array, sl[:k2], order, cores)
            sb2 = sum(s for _, s in sl[:k2])
            res["cpu_baseline"]["as_shipped_scalar"] = {
                "value": sb2 / (te2 + td2) / 1e9, "enc_gbs": sb2 / te2 / 1e9, "dec_gbs": sb2 / td2 / 1e9,
                "sample": "first %d slices, %s build, %d threads" % (k2, sc.kind, cores)}
        except Exception as e:      # pragma: no cover
            res["cpu_baseline"]["as_shipped_scalar"] = {"unavailable": str(e)}
    host.free()
    out_host.free()
    back_host.free()
    del d_in
    torch.cuda.empty_cache()
    return res


def codec_entry(name, r, world, ms_enc, ms_dec, e2e_ms):
    """The JSON shape of one codec workload (times already reduced over ranks)."""
    U, Cc = r["U"], r["C"]
    ms_step = ms_enc + ms_dec
    out = {
        "metric": METRIC, "value": world * U / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_step,
        "config": codec_config(name, r["order"], U, r["S"], r["n"], world, Cc / U),
        "enc_gbs": world * U / (ms_enc * 1e-3) / 1e9, "dec_gbs": world * U / (ms_dec * 1e-3) / 1e9,
        "enc_packed_gbs": (world * U / (r["packed_enc_ms"] * 1e-3) / 1e9) if r.get("packed_enc_ms") else None,
        "e2e": {"value": world * U / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
                "h2d_bytes_per_step": U + Cc, "d2h_bytes_per_step": U + Cc,
                "enc_gbs": U / r["e2e_enc_s"] / 1e9, "dec_gbs": U / r["e2e_dec_s"] / 1e9},
        "gpu_launches": r["gpu_launches"], "stream_flags": r["flags"],
        "roofline": r["roofline"], "roofline_enc": r["roofline_enc"], "roofline_dec": r["roofline_dec"],
        "roofline_step": r["roofline_step"], "clocks": r["clocks"], "wall_s_timed_region": r["wall_s_timed_region"],
    }
    pc = pcie_ceiling(world, Cc / U)
    if pc:
        out["e2e"]["pcie"] = pc
        out["e2e"]["frac_of_pcie_ceiling"] = out["e2e"]["value"] / pc["round_trip_ceiling_gbs"]
    for k in ("cpu_baseline", "parity_checked_slices"):
        if k in r:
            out[k] = r[k]
    return out


# ------------------------------------------------------------------ config 5: FASTQ blocks, fqzcomp5 -3 semantics
def make_fastq_block(seq, qual, first_record, nrec=BLOCK_RECORDS, rl=READ_LEN):
    """@SIM.<9 digits> <9 digits>/1 \\n seq \\n + \\n qual \\n -- fixed width, so numpy can build it."""
    head = np.frombuffer(b"@SIM.000000000 000000000/1\n", np.uint8)
    w = head.size + rl + 1 + 2 + rl + 1
    rec = np.empty((nrec, w), np.uint8)
    rec[:, :head.size] = head
    idx = np.arange(nrec, dtype=np.int64) + first_record
    for d in range(9):
        dig = ((idx // 10 ** (8 - d)) % 10 + 48).astype(np.uint8)
        rec[:, 5 + d] = dig
        rec[:, 15 + d] = dig
    o = head.size
    rec[:, o:o + rl] = seq[:nrec * rl].reshape(nrec, rl)
    rec[:, o + rl] = 10
    rec[:, o + rl + 1] = ord("+")
    rec[:, o + rl + 2] = 10
    rec[:, o + rl + 3:o + 2 * rl + 3] = (qual[:nrec * rl] + 33).reshape(nrec, rl)
    rec[:, w - 1] = 10
    return rec.reshape(-1)


def cpu_trial_sections(codec, fields, lists, S, threads, limit_bytes):
    """compress_with_methods' serial loop (fqzcomp5.c:1979-2119, rANS members) over slices of the sections on
    `threads` host threads, bounded to ~limit_bytes of section data.  Returns (seconds, bytes, results) with
    results[(sec, i)] = (sizes per method, winner index, winner bytes)."""
    jobs = []
    taken = 0
    for s, (data, methods) in enumerate(zip(fields, lists)):
        for i, o in enumerate(range(0, max(len(data), 1), S)):
            if taken >= limit_bytes and i:
                break
            jobs.append((s, i, o, min(S, len(data) - o), methods))
            taken += min(S, len(data) - o)
    res = {}
    bound = max(codec.bound(S, m) for _, ms in zip(fields, lists) for m in ms) + 64
    tl = threading.local()

    def work(j):
        s, i, o, ln, methods = jobs[j]
        if not hasattr(tl, "bufs"):
            tl.bufs = [np.empty(bound, np.uint8) for _ in range(2)]
        best, bsz, sizes = -1, None, []
        base = fields[s].ctypes.data + o
        cur = 0
        for q, m in enumerate(methods):
            sz = codec.compress_into(base, ln, tl.bufs[cur].ctypes.data, bound, m)
            sizes.append(max(sz, 0))
            if sz > 0 and (bsz is None or bsz > sz):
                best, bsz = q, sz
                cur ^= 1                     # keep the winner, reuse the other buffer
        res[(s, i)] = (sizes, best, bytes(tl.bufs[cur ^ 1][:bsz]) if best >= 0 else None)
    t = run_threads(len(jobs), work, threads)
    return t, taken, res


def run_blocks(args, env, seq, qual, ngpu):
    """configs[4]: B blocks per GPU of 1 GB FASTQ text through the block calls of the library, dealt over
    ngpu devices by its persistent workers (one process), stock 4-lane method lists and with X32."""
    bc = env["bc"]
    import ctypes as C
    nb = args.blocks_per_gpu * ngpu
    ndist = 2 if seq.size >= 2 * BLOCK_RECORDS * READ_LEN else 1
    texts = []
    for d in range(ndist):
        t = make_fastq_block(seq[d * BLOCK_RECORDS * READ_LEN:], qual[d * BLOCK_RECORDS * READ_LEN:], d * BLOCK_RECORDS)
        pb = bc.PinnedBuffer(t.size)
        pb.array[:] = t
        texts.append(pb)
    n = int(texts[0].array.size)
    # workers per device the library's block calls use: block b runs on worker (b % ngpu, (b // ngpu) % W)
    W = max(1, min(4, int(os.environ.get("B200RANS_BLOCK_WORKERS", "2"))))
    out_cap = n // 2 + (64 << 20)
    outs = [bc.PinnedBuffer(out_cap) for _ in range(ngpu * W)]
    backs = [bc.PinnedBuffer(n + 4096) for _ in range(ngpu * W)]
    slot = lambda b: partition.worker_slot(b, ngpu, W)
    tlist = [texts[b % ndist].array for b in range(nb)]
    olist = [outs[slot(b)].array for b in range(nb)]
    blist = [backs[slot(b)].array for b in range(nb)]
    variants = {}
    S = 262144
    for vname, x32 in (("stock_4lane", False), ("x32", True)):
        opts = bc.block_opts(slice_bytes=S, seq=M3_SEQ, qual=M3_QUAL, names=M3_NAMES, x32=x32)
        launches0 = bc.launch_count()
        te, td = [], []
        reps = dreps = None
        for it in range(1 + max(1, min(args.steps, 2))):
            t0 = time.perf_counter()
            reps = bc.encode_blocks_multi(ngpu, tlist, opts, olist)
            t1 = time.perf_counter()
            assert all(r.status == 0 for r in reps), "a block failed to encode"
            # decode the blocks that are still in the output buffers: the last one each worker wrote
            last = {}
            for b in range(nb):
                last[slot(b)] = b
            live = sorted(last.values())
            t2 = time.perf_counter()
            dreps = bc.decode_blocks_multi(ngpu, [olist[b] for b in live], [reps[b].block_len for b in live],
                                           [blist[b] for b in live])
            t3 = time.perf_counter()
            assert all(r.status == 0 for r in dreps), "a block failed to decode"
            if it:
                te.append(t1 - t0)
                td.append((t3 - t2) * nb / len(live))      # scaled to nb blocks
        launches = bc.launch_count() - launches0
        for j, b in enumerate(live):
            assert dreps[j].block_len == n and np.array_equal(blist[b][:n], tlist[b]), "block round trip mismatch"
        r0 = reps[0]
        te_m, td_m = float(np.mean(te)), float(np.mean(td))
        variants[vname] = {
            "value": nb * n / (te_m + td_m) / 1e9, "unit": "GB/s of FASTQ text, round trip, host buffers",
            "enc_gbs": nb * n / te_m / 1e9, "dec_gbs": nb * n / td_m / 1e9,
            "enc_s": te_m, "dec_s_scaled": td_m, "blocks": nb, "blocks_decoded_per_pass": len(live),
            "block_bytes": n, "block_out_bytes": int(r0.block_len), "ratio": r0.block_len / n,
            "slices": [int(x) for x in r0.nslices], "gpu_launches": int(launches),
            "methods": {"names": list(M3_NAMES), "seq": [m | (4 if x32 else 0) for m in M3_SEQ],
                        "qual": ["(150<<8)+9" if m < 0 else m | (4 if x32 else 0) for m in M3_QUAL]},
            "wins": {"seq": [int(x) for x in list(r0.wins[1])[:len(M3_SEQ)]],
                     "qual": [int(x) for x in list(r0.wins[2])[:len(M3_QUAL)]]},
            "encode_phase_ms_per_block": {k: float(np.mean([r.ms[i] for r in reps])) for i, k in enumerate(
                ("copy_in_split", "plan_and_queue", "wait_trials", "frame_crc_copy_out"))},
        }
        variants[vname]["_last"] = (reps, live)
        # ---- the same blocks as fqzcomp5 -3 really codes them: its learner (metrics_method / metrics_update,
        # fqzcomp5.c:1899-1958) tries every method on METRICS_TRIAL = 3 blocks, then uses the best one alone for
        # the next METRICS_REVIEW = 100
        tl, tdl, picked = [], [], None
        nbl = args.blocks_total                      # the named config: 64 blocks, whatever the number of GPUs
        tlist_l = [texts[b % ndist].array for b in range(nbl)]
        olist_l = [outs[slot(b)].array for b in range(nbl)]
        blist_l = [backs[slot(b)].array for b in range(nbl)]
        for it in range(1 + max(1, min(args.steps, 2))):
            L = bc.Learner()
            t0 = time.perf_counter()
            ntr = min(3, nbl)
            o_tr = [L.methods(opts) for _ in range(ntr)]
            reps_a = bc.encode_blocks_multi(ngpu, tlist_l[:ntr], o_tr[0], olist_l[:ntr])
            for o, r in zip(o_tr, reps_a):
                L.update(o, r)
            o_st = [L.methods(opts) for _ in range(nbl - ntr)]
            reps_b = bc.encode_blocks_multi(ngpu, tlist_l[ntr:], o_st[0], olist_l[:nbl - ntr]) if nbl > ntr else []
            t1 = time.perf_counter()
            assert all(r.status == 0 for r in list(reps_a) + list(reps_b)), "a block failed to encode"
            m = nbl - ntr
            last = {}
            for b in range(m):
                last[slot(b)] = b
            live = sorted(last.values())
            t2 = time.perf_counter()
            dr = bc.decode_blocks_multi(ngpu, [olist_l[b] for b in live], [reps_b[b].block_len for b in live],
                                        [blist_l[b] for b in live]) if m else []
            t3 = time.perf_counter()
            assert all(r.status == 0 for r in dr), "a block failed to decode"
            for j, b in enumerate(live):
                assert dr[j].block_len == n and np.array_equal(blist_l[b][:n], tlist_l[ntr + b]), "block round trip mismatch"
            if it:
                tl.append(t1 - t0)
                tdl.append((t3 - t2) * nbl / max(len(live), 1))
            if m:
                picked = {"seq": int(o_st[0].seq_methods[0]), "qual": int(o_st[0].qual_methods[0]),
                          "n_seq": int(o_st[0].n_seq_methods), "n_qual": int(o_st[0].n_qual_methods),
                          "steady_block_out_bytes": int(reps_b[0].block_len),
                          "steady_phase_ms_per_block": {k: float(np.mean([r.ms[i] for r in reps_b])) for i, k in
                                                        enumerate(("copy_in_split", "plan_and_queue", "wait_trials",
                                                                   "frame_crc_copy_out"))}}
        tl_m, tdl_m = float(np.mean(tl)), float(np.mean(tdl))
        variants[vname + "_learner"] = {
            "value": nbl * n / (tl_m + tdl_m) / 1e9, "unit": "GB/s of FASTQ text, round trip, host buffers",
            "enc_gbs": nbl * n / tl_m / 1e9, "dec_gbs": nbl * n / tdl_m / 1e9, "enc_s": tl_m, "dec_s_scaled": tdl_m,
            "blocks": nbl, "trial_blocks": min(3, nbl), "block_bytes": n, "picked": picked, "scaling": "strong",
            "note": "fqzcomp5's learner: METRICS_TRIAL = 3 blocks try every method, the rest use the best one alone"}
    # ---- parity on the timed data + CPU codec-only baseline: the same serial trial loop on the host
    cores = os.cpu_count() or 1
    res = {"metric": "fqzcomp5 -3 block pipeline GB/s of FASTQ text (split + method trials + framing + CRC; "
                     "CRC check + decode + join), configs[4]",
           "config": {"workload": BLOCKS, "blocks": args.blocks_total, "blocks_all_on_trial_variants": nb,
                      "blocks_per_gpu_all_on_trial_variants": args.blocks_per_gpu, "n_gpus": ngpu,
                      "block_bytes": n, "distinct_blocks": ndist, "records_per_block": BLOCK_RECORDS,
                      "slice_bytes": S, "dispatch": "one process, %d persistent workers per device, block b on "
                      "device b %% ngpu, results in dispatch order" % W,
                      "note": "blocks cycle over %d distinct synthetic texts held in pinned host memory; LZP and "
                              "tok3 (host stages of -3) are outside the path, names go through rANS order 5" % ndist},
           "unit": "GB/s"}
    from oracle.pyoracle import FastqChecker, fastq_available
    fqc = FastqChecker("ref" if fastq_available("ref") else "oracle")
    small = bytes(texts[0].array[:331 * 40000])                  # 40 000 records: what the CPU splits in a moment
    want = fqc.split(small)
    codec = cpu_codec()
    checked = 0
    for vname, x32 in (("stock_4lane", False), ("x32", True)):
        opts = bc.block_opts(slice_bytes=S, seq=M3_SEQ, qual=M3_QUAL, names=M3_NAMES, x32=x32)
        blk, rep = bc.encode_block(small, opts)
        assert rep.status == 0 and rep.num_records == want["num_records"]
        fields = [np.frombuffer(want[k], np.uint8) for k in ("name", "seq", "qual")]
        lists = [bc.resolve_methods(l, want["fixed_len"], x32) for l in (M3_NAMES, M3_SEQ, M3_QUAL)]
        Ss = [S, S - S % want["fixed_len"], S - S % want["fixed_len"]]
        b = bytes(blk)
        p = 12
        for s in range(3):
            clen = int.from_bytes(b[p + 5:p + 9], "little")
            q = p + 9
            nsl = int.from_bytes(b[q:q + 4], "little")
            cs = np.frombuffer(b[q + 8:q + 8 + 4 * nsl], np.uint32)
            q += 8 + 4 * nsl
            _, _, cres = cpu_trial_sections(codec, [fields[s]], [lists[s]], Ss[s], cores, 1 << 40)
            sums = [0] * len(lists[s])
            for i in range(nsl):
                sizes, best, wbytes = cres[(0, i)]
                assert b[q:q + int(cs[i])] == wbytes, "block slice differs from the CPU trial's winner"
                sums = [a + c for a, c in zip(sums, sizes)]
                q += int(cs[i])
                checked += 1
            assert list(rep.csize[s])[:len(sums)] == sums
            p += 9 + clen
            if s == 0:
                p += 1 + b[p]
    res["parity_checked_slices"] = checked
    if ngpu == 1 and not args.no_cpu:
        # CPU codec-only: the serial trial loop over section slices of one block, bounded
        full = fqc.split(bytes(texts[0].array[:331 * 400000]))
        fields = [np.frombuffer(full[k], np.uint8) for k in ("name", "seq", "qual")]
        cpu = {}
        for vname, x32 in (("stock_4lane", False), ("x32", True)):
            lists = [bc.resolve_methods(l, full["fixed_len"], x32) for l in (M3_NAMES, M3_SEQ, M3_QUAL)]
            t, by, _ = cpu_trial_sections(codec, fields, lists, S - S % 150, cores, 48 << 20)
            cpu[vname] = {"value": by / t / 1e9, "unit": "GB/s of section data (encode side only)", "cores": cores,
                          "kind": kind_name(codec), "sample": "%.0f MB of one block's sections, %s build, %d "
                          "threads, the serial method loop per slice" % (by / 1e6, codec.kind, cores)}
        res["cpu_baseline"] = {"codec_only": cpu}
        res["cpu_baseline"]["e2e_tool"] = tool_baseline(texts, cores)
    for v in variants.values():
        v.pop("_last", None)
    res["variants"] = variants
    res["value"] = variants["x32_learner"]["value"]
    res["value_is"] = "x32_learner"
    vx, vl = variants["x32"], variants["x32_learner"]
    ob = int((vl["picked"] or {}).get("steady_block_out_bytes", vx["block_out_bytes"]))
    res["e2e"] = {"value": vl["value"], "unit": "GB/s", "all_methods_every_block": vx["value"],
                  "h2d_bytes_per_step": vl["blocks"] * n, "d2h_bytes_per_step": vl["blocks"] * ob,
                  "note": "the block calls take and return host buffers: this IS the end-to-end number (bytes: "
                          "the encode pass over all blocks; the decode pass moves the same the other way)"}
    peak, peak_src = load_peaks()
    a = vl["blocks"] * (n + ob) / vl["enc_s"] / 1e9 / max(ngpu, 1)
    res["roofline"] = {"kernel": "whole encode chain per GPU (host buffers in and out)", "bound": "hbm",
                       "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None,
                       "peak_source": peak_src,
                       "note": "text in + block out per second and GPU; steady-state blocks are bound by the PCIe "
                               "copy of the text (see encode_phase_ms), trial blocks by the STRIPE candidate's "
                               "300 sub-streams per slice"}
    for t in texts + outs + backs:
        t.free()
    return res


def tool_baseline(texts, cores):
    """The reference tool itself, `fqzcomp5 -3 -t nproc`, on the config's FASTQ written to a RAM-backed file:
    wall time of compress and of decompress (SURVEY 8d CPU baseline (2))."""
    exe = os.path.join(ROOT, "oracle", "_ref", "fqzcomp5_ref")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/fqzcomp5_ref not built"}
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None)
    try:
        fq = os.path.join(d, "blocks.fq")
        total = 0
        with open(fq, "wb") as f:
            for t in texts:
                f.write(memoryview(t.array))
                total += int(t.array.size)
        out = {}
        for label, bs in (("-b 1G", "1G"), ("-b 100M", "100M")):
            fz, back = os.path.join(d, "o.fqz5"), os.path.join(d, "o.fq")
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-3", "-b", bs, "-t", str(cores), fq, fz], capture_output=True, timeout=900)
            t1 = time.perf_counter()
            if r.returncode:
                out[label] = {"unavailable": "fqzcomp5_ref exit %d" % r.returncode}
                continue
            r = subprocess.run([exe, "-d", "-t", str(cores), fz, back], capture_output=True, timeout=900)
            t2 = time.perf_counter()
            ok = r.returncode == 0 and os.path.getsize(back) == total
            out[label] = {"enc_gbs": total / (t1 - t0) / 1e9, "dec_gbs": total / (t2 - t1) / 1e9 if ok else None,
                          "value": total / (t2 - t0) / 1e9 if ok else None, "unit": "GB/s of FASTQ text",
                          "out_bytes": os.path.getsize(fz), "blocks": -(-total // (10 ** 9 if bs == "1G" else 10 ** 8)),
                          "roundtrip_ok": bool(ok)}
            os.remove(fz)
            if os.path.exists(back):
                os.remove(back)
        out["cores"] = cores
        out["kind"] = "reference"
        out["sample"] = ("fqzcomp5_ref -3 -t %d on %.1f GB of the same FASTQ text (a RAM-backed file); the tool "
                         "parallelises over blocks only, so -b 1G keeps %d threads busy and -b 100M (the -3 "
                         "default) all of them; it also runs LZP and tok3, which the GPU path leaves on the host"
                         % (cores, total / 1e9, len(texts)))
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ------------------------------------------------------------------ GPU arm: driver
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    names = [args.workload] if args.workload != "all" else [HEADLINE] + (PER_CONFIG if world == 1 else [])
    want_blocks = args.workload in ("all", BLOCKS)
    codec_names = [w for w in names if w in WORKLOADS]
    # ---- synthetic inputs first (forked generator processes, before CUDA is initialised)
    cores = os.cpu_count() or 1
    synth.PROCESSES = max(1, min(cores // max(world, 1), 32))
    data = {}
    for w in codec_names:
        gen, _, total, _ = WORKLOADS[w]
        if args.bytes:
            total = args.bytes
        if (gen, total) not in data:
            data[(gen, total)] = synth.GENERATORS[gen](total, seed=synth_seed(gen) + 7919 * rank)
    bseq = bqual = None
    if want_blocks and rank == 0:
        need = 2 * BLOCK_RECORDS * READ_LEN
        q = data.get(("illumina_qual", 999_999_900))
        s = data.get(("illumina_seq", 999_999_900))
        bqual = q if q is not None else synth.illumina_qual(need, seed=synth_seed("illumina_qual"))
        bseq = s if s is not None else synth.illumina_seq(need, seed=synth_seed("illumina_seq"))
    synth.PROCESSES = 0

    from fqzcomp5_b200 import codec as bc
    torch.cuda.set_device(local)
    bc.lib().b200rans_set_device(local)
    if world > 1:
        dist.init_process_group("gloo")      # barrier and max-over-ranks only; no collective on the data path
    dev = torch.device("cuda", local)
    tstream = torch.cuda.Stream(device=dev)   # a dedicated stream: the library launches on it, events are recorded on it
    torch.cuda.set_stream(tstream)
    assert tstream.cuda_stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    env = {"torch": torch, "bc": bc, "dev": dev, "stream": tstream.cuda_stream, "local": local, "barrier": barrier}

    def reduce_max(vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    entries = {}
    for w in codec_names:
        gen, _, total, _ = WORKLOADS[w]
        if args.bytes:
            total = args.bytes
        r = measure_codec(args, w, data[(gen, total)], env, with_cpu=(world == 1 and not args.no_cpu))
        ms_enc, ms_dec, e2e_ms = reduce_max([r["ms_enc"], r["ms_dec"], (r["e2e_enc_s"] + r["e2e_dec_s"]) * 1e3])
        entries[w] = codec_entry(w, r, world, ms_enc, ms_dec, e2e_ms)
    blocks = None
    if want_blocks:
        barrier()
        if rank == 0:
            blocks = run_blocks(args, env, bseq, bqual, world)
        barrier()
    if rank == 0:
        if codec_names:
            head = entries[codec_names[0]]
            line = {"metric": METRIC, "value": head["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic"}
            line.update({k: v for k, v in head.items() if k not in ("metric", "value", "unit", "ms_per_step")})
            pc = {w: entries[w] for w in codec_names[1:]}
            if blocks is not None:
                pc[BLOCKS] = blocks
            if pc:
                line["per_config"] = pc
        else:
            line = {"metric": blocks["metric"], "value": blocks["value"], "unit": "GB/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": (blocks["variants"]["x32"]["enc_s"] + blocks["variants"]["x32"]["dec_s_scaled"]) * 1e3,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32",
                    "data": "synthetic"}
            line.update({k: v for k, v in blocks.items() if k not in ("metric", "value", "unit")})
            line["gpu_launches"] = blocks["variants"]["x32"]["gpu_launches"]
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


class QuietStdout:
    """Route everything libraries print on fd 1 to stderr, so that stdout carries exactly one JSON line;
    emit() writes to the real stdout."""

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
    ap.add_argument("--workload", default="all", choices=["all", BLOCKS] + sorted(WORKLOADS) + sorted(EXTRA))
    ap.add_argument("--bytes", type=int, default=0, help="block size per GPU (default: the config's 1 GB)")
    ap.add_argument("--slice", type=int, default=256 << 10, help="bytes per rans_compress_to_4x16 call")
    ap.add_argument("--blocks-per-gpu", type=int, default=8,
                    help="fastq_blocks_m3, every block on trial: blocks per GPU (8 x 8 GPUs = 64)")
    ap.add_argument("--blocks-total", type=int, default=64,
                    help="fastq_blocks_m3 with fqzcomp5's learner: blocks in all, at any number of GPUs (configs[4]: 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    if args.workload in EXTRA:
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                emit({"impl": "reference", "unavailable": "the CPU leg of %s is reported inside the b200 arm's "
                      "cpu_baseline (one thread: the reference runs this step on one thread)" % args.workload})
            return
        return run_extra(args)
    if args.impl == "reference":
        name = HEADLINE if args.workload in ("all", BLOCKS) else args.workload
        run_reference(args, name)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
