"""One device-resident encode (in-slot) and one decode of a bench workload at its full size: the launches of this
process under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` are exactly
one encode step and one decode step (scripts/traffic_from_ncu.py turns the CSVs into profiles/traffic.json).
usage: traffic_probe.py <workload> [bytes]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                        # noqa: E402
from fqzcomp5_b200 import synth, codec as bc        # noqa: E402

name = sys.argv[1]
gen, order, total, _ = bench.WORKLOADS[name]
if len(sys.argv) > 2:
    total = int(sys.argv[2])
S = 256 << 10
synth.PROCESSES = min(os.cpu_count() or 1, 16)
data = synth.GENERATORS[gen](total, seed=bench.synth_seed(gen))
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
sl = synth.slices(data, S)
n = len(sl)
in_off = np.array([o for o, _ in sl], np.uint64)
in_size = np.array([s for _, s in sl], np.uint32)
orders = np.full(n, order, np.int32)
d_in = torch.from_numpy(data).to(dev)
cap = bc.compress_slots_bound(in_size, orders)
d_comp = torch.empty(cap, dtype=torch.uint8, device=dev)
d_coff = torch.zeros(n, dtype=torch.int64, device=dev)
d_csz = torch.zeros(n, dtype=torch.int32, device=dev)
d_back = torch.empty(total, dtype=torch.uint8, device=dev)
d_osz = torch.zeros(n, dtype=torch.int32, device=dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
bc.compress_batch_dev2(stream, d_in.data_ptr(), in_off, in_size, orders, d_comp.data_ptr(), cap,
                       d_coff.data_ptr(), d_csz.data_ptr(), flags=bc.OUT_IN_SLOT)
torch.cuda.synchronize()
coff = d_coff.cpu().numpy().astype(np.uint64)
csz = d_csz.cpu().numpy().astype(np.uint32)
flags = d_comp[torch.from_numpy(coff.astype(np.int64)).to(dev)].cpu().numpy()
bc.uncompress_batch_dev(stream, d_comp.data_ptr(), coff, csz, d_back.data_ptr(), in_off, in_size,
                        d_osz.data_ptr(), d_st.data_ptr(), flags=flags)
torch.cuda.synchronize()
assert int(d_st.abs().sum()) == 0 and torch.equal(d_back, d_in)
print("workload %s: U %d C %d streams %d" % (name, total, int(csz.sum()), n))
