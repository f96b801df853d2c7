"""Differential test aimed at rans_compute_shift's decision boundary (rANS_static4x16pr.c:357-420).

The order-1 encoder chooses 10 or 12 bits of precision from two entropy estimates in double precision:
shift = 10 iff e10 / e12 < 1.01 (or no context needs more than 1024).  The library takes that decision
on the device (CUDA's log(), per-row sums reduced across warps), the reference on the host under
-ffast-math, so inputs whose ratio sits at the threshold are where the two could part.  For a family of
seeded sources this script finds, by bisection on the input LENGTH, prefixes whose ratio is within
`tol` of 1.01 (the estimate is recomputed here in numpy only to steer the search), then compares GPU
and reference on them: bytes equal, the first table byte (shift << 4) equal, and the round trip.

usage: gpu_fuzz_shift.py [sources] [seed] [tol]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fqzcomp5_b200 import codec                     # noqa: E402
from oracle.pyoracle import Codec, available        # noqa: E402


def shift_ratio(d, N=32):
    """e10 / e12 and max_tot as encode_freq1 + rans_compute_shift see them (rANS_static16_int.h:312-366)."""
    n = d.size
    F = np.zeros((256, 256), np.int64)
    prev = np.empty(n, np.int64)
    prev[0] = 0
    prev[1:] = d[:-1]
    np.add.at(F, (prev, d.astype(np.int64)), 1)
    T = F.sum(1)
    T[d[-1]] += 1                                    # utils.h:311,345
    seg = n // N
    for z in range(1, N):                            # :325-327
        F[0, d[z * seg]] += 1
    T[0] += N - 1
    e10 = e12 = 0.0
    max_tot = 0
    F0 = np.bincount(d, minlength=256)
    F0[0] = max(F0[0], 1)
    flog = lambda a: (np.asarray(a, np.float64).view(np.int64) - 4606921278410026770) * 1.539095918623324e-16
    for i in np.nonzero(F0)[0]:
        if T[i] == 0:
            continue
        mv = 1 << int(T[i] - 1).bit_length() if T[i] > 1 else 1
        f = F[i][F[i] > 0].astype(np.float64)
        sm10 = int((mv // F[i][F[i] > 0] > 1024).sum())
        sm12 = int((mv // F[i][F[i] > 0] > 4096).sum())
        l10, l12 = np.log(1024 + sm10), np.log(4096 + sm12)
        e10 += float(-(f * (flog(np.maximum(f * (1024.0 / T[i]), 1)) - l10)).sum() + 1.3 * f.size)
        e12 += float(-(f * (flog(np.maximum(f * (4096.0 / T[i]), 1)) - l12)).sum() + 4.7 * f.size)
        if f.size < 64 and mv > 128:
            mv //= 2
        if mv > 1024:
            mv //= 2
        mv = min(mv, 4096)
        max_tot = max(max_tot, mv)
    return e10 / e12, max_tot


def source(rng, n):
    """A long seeded stream whose order-1 statistics are skewed enough for 12 bits to pay at some length."""
    kind = int(rng.integers(0, 3))
    if kind == 0:        # sticky symbols with rare excursions (quality-like)
        k = int(rng.integers(8, 60))
        stay = rng.uniform(0.85, 0.995)
        v = rng.integers(0, k, n).astype(np.uint8)
        keep = rng.random(n) < stay
        keep[0] = False
        idx = np.where(~keep, np.arange(n), 0)
        np.maximum.accumulate(idx, out=idx)
        return (v[idx] + int(rng.integers(0, 150))).astype(np.uint8)
    if kind == 1:        # one dominant symbol per context, long-tailed others
        k = int(rng.integers(20, 200))
        p = rng.random(k) ** rng.uniform(3, 8)
        p /= p.sum()
        return rng.permutation(256)[:k].astype(np.uint8)[rng.choice(k, n, p=p)]
    walk = np.cumsum(rng.choice(np.array([-1, 0, 0, 0, 0, 0, 0, 1]), n))       # slow random walk
    return (np.clip(walk - walk.min(), 0, 200) % int(rng.integers(30, 200))).astype(np.uint8)


def boundary_lengths(d, tol, lo=2000, hi=None):
    """Lengths n (prefixes of d) with |e10/e12 - 1.01| <= tol and max_tot > 1024, found by bisection."""
    hi = hi or d.size
    r_lo, _ = shift_ratio(d[:lo])
    r_hi, mt = shift_ratio(d[:hi])
    if not (r_lo < 1.01 <= r_hi) or mt <= 1024:
        return []
    while hi - lo > 1:
        mid = (lo + hi) // 2
        r, _ = shift_ratio(d[:mid])
        if r < 1.01:
            lo = mid
        else:
            hi = mid
    out = []
    for n in range(max(2000, lo - 6), lo + 8):
        r, mt = shift_ratio(d[:n])
        if abs(r - 1.01) <= tol and mt > 1024:
            out.append((n, r))
    return out


def main(argv):
    nsrc = int(argv[1]) if len(argv) > 1 else 6
    rng = np.random.default_rng(int(argv[2]) if len(argv) > 2 else 2024)
    tol = float(argv[3]) if len(argv) > 3 else 1e-3
    ref = Codec("ref") if available("ref") else Codec("oracle")
    cases = bad = tried = 0
    shifts = {10: 0, 12: 0}
    while cases < nsrc * 8 and tried < nsrc * 40:
        tried += 1
        d = source(rng, 400000)
        for n, r in boundary_lengths(d, tol):
            data = d[:n].tobytes()
            for order in (5, 1):
                want = ref.compress(data, order)
                got = codec.rans_compress_to_4x16(data, order)
                cases += 1
                if want is None or want[0] & 0x20:
                    continue
                hdr = 1 + (0 if want[0] & 0x10 else next(i for i in range(1, 6) if not want[i] & 0x80))
                shifts[want[hdr] >> 4] = shifts.get(want[hdr] >> 4, 0) + 1
                if want != got:
                    bad += 1
                    print("ENC", n, hex(order), "ratio %.6f" % r, "ref shift", want[hdr] >> 4,
                          "gpu shift", got and got[hdr] >> 4)
                    continue
                if codec.rans_uncompress_4x16(want) != data:
                    bad += 1
                    print("DEC", n, hex(order))
    print("boundary cases", cases, "mismatches", bad, "shifts", shifts, "tol", tol)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv) else 0)
