"""Host-side logic that needs no GPU: generators, slicing, and the N>1 block
partition exercised with two real processes over gloo."""
import os

import numpy as np
import pytest

from fqzcomp5_b200 import partition, synth


def test_generators_are_deterministic_and_shaped():
    a, b = synth.illumina_qual(100000), synth.illumina_qual(100000)
    assert np.array_equal(a, b) and a.min() >= 2 and a.max() <= 40
    s = synth.illumina_seq(30000)
    assert set(np.unique(s).tolist()) <= set(b"ACGT")
    o = synth.ont_qual(100000)
    assert o.min() >= 1 and o.max() <= 40 and o.size == 100000
    bq = synth.binned_qual(5000)
    assert set(np.unique(bq).tolist()) <= {2, 12, 23, 37}
    big = synth.illumina_qual((1 << 25) + 12345)          # crosses a chunk boundary
    assert big.size == (1 << 25) + 12345


def test_slices_cover_the_buffer():
    buf = np.zeros(1000003, np.uint8)
    sl = synth.slices(buf, 262144)
    assert sl[0] == (0, 262144) and sum(s for _, s in sl) == buf.size
    assert all(o2 == o1 + s1 for (o1, s1), (o2, _) in zip(sl, sl[1:]))


def test_partition_round_robin():
    for world in (1, 2, 4, 8):
        owned = [partition.blocks_of_rank(64, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(64))
        per_rank = [[("blk", b) for b in owned[r]] for r in range(world)]
        assert partition.gather_in_order(per_rank, 64, world) == [("blk", b) for b in range(64)]


def test_worker_slots_never_collide_in_flight():
    """bench.py gives every (device, worker) of the library's block calls its own output buffer: block b runs on
    worker (b % ngpu, (b // ngpu) % W), so blocks sharing a slot are W * ngpu dispatches apart."""
    for ngpu in (1, 2, 4, 8):
        for W in (1, 2, 3):
            slots = [partition.worker_slot(b, ngpu, W) for b in range(64)]
            assert set(slots) == set(range(ngpu * W))
            for b in range(64 - ngpu * W):
                assert len(set(slots[b:b + ngpu * W])) == ngpu * W
            assert all(s // W == partition.owner(b, ngpu) for b, s in enumerate(slots))


def _worker(rank, world, port, nblocks, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = partition.blocks_of_rank(nblocks, rank, world)
    # stand-in for the per-block compressed size each rank would produce
    sizes = torch.tensor([1000 + 7 * b for b in mine] + [0] * (nblocks - len(mine)), dtype=torch.int64)
    gathered = [torch.zeros(nblocks, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, sizes)                       # index/size exchange only: no data-path collective
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)               # the max-over-ranks timing reduction bench.py uses
    if rank == 0:
        per_rank = [gathered[r][:len(partition.blocks_of_rank(nblocks, r, world))].tolist() for r in range(world)]
        q.put((partition.gather_in_order(per_rank, nblocks, world), float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, nblocks, port = 2, 7, 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, nblocks, q)) for r in range(world)]
    for p in ps:
        p.start()
    res, tmax = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [1000 + 7 * b for b in range(nblocks)] and tmax == 2.0


def test_tok3_method_lists_follow_the_reference_filters():
    """tokenise_name3.c:1268-1417: levels 1-9 map to table rows 0-4; the rANS build clears X32 (0x04)
    from every entry; STRIPE entries are skipped unless the stream length is a multiple of 4."""
    from fqzcomp5_b200 import codec
    assert [codec.tok3_level_row(l) for l in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12)] == [0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4]
    assert len(codec.TOK3_METHODS) == 5 and all(len(r) == len(codec.TOK3_TYPES) == 13 for r in codec.TOK3_METHODS)
    row9 = codec.TOK3_METHODS[4]
    digits = row9[codec.TOK3_TYPES.index("DIGITS")]                  # {132, 201, 1, 192, 129, 193}
    assert codec.tok3_method_list(digits, 4000) == [128, 201, 1, 192, 129, 193]      # 132 -> 128 (X32 cleared)
    assert codec.tok3_method_list(digits, 4001) == [128, 1, 192, 129, 193]           # 201 has STRIPE: skipped
    assert codec.tok3_method_list(codec.TOK3_METHODS[0][codec.TOK3_TYPES.index("DUP")], 7) == []   # {8}, len % 4
    assert codec.ransxn1_order(150) == (150 << 8) + 9                                # fqzcomp5.c:2019


def test_bench_extra_workloads_are_registered():
    import bench
    assert set(bench.EXTRA) == {"fastq_split", "fastq_join", "crc32"}
    assert "illumina_qual_o0" in bench.WORKLOADS and bench.METRIC.startswith("rANS32x16")


def test_tok3_tables_in_c():
    """b200rans_tok3_methods (C, no device needed) against the Python transcription of
    tokenise_name3.c:1283-1357 with the filters of :1374-1378."""
    from fqzcomp5_b200 import codec
    for level in range(1, 10):
        row = codec.TOK3_METHODS[codec.tok3_level_row(level)]
        for t in range(13):
            for n in (0, 3, 4, 1001):
                assert codec.tok3_methods(level, t, n) == codec.tok3_method_list(row[t], n)
    import pytest
    with pytest.raises(codec.B200RansError):
        codec.tok3_methods(3, 13, 4)


def test_method_learner_follows_metrics_method():
    """b200fqz_learner_* (host logic, no device) against a restatement of metrics_method / metrics_update
    (fqzcomp5.c:1899-1958): METRICS_TRIAL = 3 blocks with every method, the best (csize + 1) / usize alone for
    METRICS_REVIEW = 100 blocks, then the next trial; a skipped RANSXN1 leaves a hole in the report's sizes."""
    import numpy as np
    from fqzcomp5_b200 import codec
    rng = np.random.default_rng(5)
    allo = codec.block_opts(slice_bytes=262144, x32=True)
    lists = [[5], [4, 5, 133, 197], [4, 5, 133, 197, codec.RANSXN1]]

    class Ref:                                          # the reference's per-section state machine
        def __init__(self, n):
            self.n, self.review, self.trial, self.used = n, 0, -99999, 0
            self.us, self.cs = [0] * n, [0] * n

        def methods(self):
            if self.n <= 1:
                return list(range(self.n)), False
            if self.review <= 0:
                self.review, self.trial = 100, 3
                self.us, self.cs = [0] * self.n, [0] * self.n
            if self.trial > 0:
                return list(range(self.n)), True
            if self.trial > -99999:
                best, bsz = 0, 1e30
                for m in range(self.n):
                    if self.us[m] and bsz > (self.cs[m] + 1.0) / self.us[m]:
                        bsz, best = (self.cs[m] + 1.0) / self.us[m], m
                self.used, self.trial = best, -99999
            else:
                self.review -= 1
            return [self.used], False

    L = codec.Learner()
    refs = [Ref(len(l)) for l in lists]
    for blk in range(230):
        o = L.methods(allo)
        fixed = 150 if blk % 7 else 0                   # every seventh block has reads of different lengths
        rep = codec.BlockReport()
        rep.status, rep.fixed_len = 0, fixed
        got = [[o.name_methods[i] for i in range(o.n_name_methods)], [o.seq_methods[i] for i in range(o.n_seq_methods)],
               [o.qual_methods[i] for i in range(o.n_qual_methods)]]
        for s in range(3):
            idx, trial = refs[s].methods()
            want = [lists[s][i] | (0 if lists[s][i] < 0 else 0) for i in idx]
            assert got[s] == want, (blk, s, got[s], want)
            rep.ulen[s] = 1000 + s
            k = 0
            for i in idx:
                if lists[s][i] == codec.RANSXN1 and fixed <= 0:
                    continue                            # skipped by the block call: no size in the report
                c = int(rng.integers(100, 900))
                rep.csize[s][k] = c
                k += 1
                if trial:
                    refs[s].us[i] += 1000 + s
                    refs[s].cs[i] += c
            if trial:
                refs[s].trial -= 1
        L.update(o, rep)
