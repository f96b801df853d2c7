"""PCIe copy rates of the box (pinned memory) with 1, 2, 4, ... GPUs copying at the same time: H2D alone,
D2H alone, both at once.  The aggregate over GPUs is the ceiling of every host-buffer (`e2e`) number and
of the multi-GPU block pipeline; bench.py's scaling of `e2e` is read against it.

usage: pcie_probe.py [max_gpus]      (one JSON line per GPU count)
"""
import json
import sys
import time

import torch

n = 1 << 30
ngpu_max = min(int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count(), torch.cuda.device_count())


class Dev:
    def __init__(self, g):
        self.g = g
        with torch.cuda.device(g):
            self.h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
            self.h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
            self.d1 = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % g)
            self.d2 = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % g)
            self.s1, self.s2 = torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)

    def h2d(self):
        with torch.cuda.stream(self.s1):
            self.d1.copy_(self.h1, non_blocking=True)

    def d2h(self):
        with torch.cuda.stream(self.s2):
            self.h2.copy_(self.d2, non_blocking=True)


def sync(devs):
    for d in devs:
        torch.cuda.synchronize(d.g)


def run(devs, what, reps=5):
    def once():
        for d in devs:
            if what in ("h2d", "both"):
                d.h2d()
            if what in ("d2h", "both"):
                d.d2h()
    once()
    sync(devs)
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    sync(devs)
    return (time.perf_counter() - t0) / reps


devs = []
g = 1
while g <= ngpu_max:
    while len(devs) < g:
        devs.append(Dev(len(devs)))
    t_h, t_d, t_b = run(devs, "h2d"), run(devs, "d2h"), run(devs, "both")
    print(json.dumps({"gpus": g, "h2d_gbs_total": g * n / t_h / 1e9, "d2h_gbs_total": g * n / t_d / 1e9,
                      "both_gbs_each_direction_total": g * n / t_b / 1e9,
                      "h2d_gbs_per_gpu": n / t_h / 1e9, "d2h_gbs_per_gpu": n / t_d / 1e9,
                      "both_gbs_each_direction_per_gpu": n / t_b / 1e9}), flush=True)
    g *= 2
