"""One fastq_blocks_m3 block through b200fqz_encode_block / decode_block (for launch lists under ncu).
usage: block_once.py [x32=1] [repeats=2]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                        # noqa: E402
from fqzcomp5_b200 import synth, codec as bc        # noqa: E402

x32 = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
synth.PROCESSES = min(os.cpu_count() or 1, 16)
need = bench.BLOCK_RECORDS * bench.READ_LEN
text = bench.make_fastq_block(synth.illumina_seq(need, seed=3), synth.illumina_qual(need, seed=2), 0)
synth.PROCESSES = 0
t = bc.PinnedBuffer(text.size)
t.array[:] = text
out = bc.PinnedBuffer(text.size // 2 + (64 << 20))
back = bc.PinnedBuffer(text.size + 4096)
opts = bc.block_opts(slice_bytes=262144, seq=bench.M3_SEQ, qual=bench.M3_QUAL, names=bench.M3_NAMES, x32=x32)
for it in range(reps):
    t0 = time.perf_counter()
    blk, rep = bc.encode_block(t.array, opts, out=out.array)
    t1 = time.perf_counter()
    txt, drep = bc.decode_block(blk, back.array.size, out=back.array)
    t2 = time.perf_counter()
    assert rep.status == 0 and drep.status == 0 and np.array_equal(txt, t.array)
    print("pass %d: encode %.1f ms (phases %s) decode %.1f ms, block %d B" % (
        it, (t1 - t0) * 1e3, [round(x, 1) for x in rep.ms], (t2 - t1) * 1e3, rep.block_len), flush=True)
