"""Per-kernel: group SASS instructions into contiguous regions by executed count and show the biggest."""
import csv, io, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = None; hdr = None; data = {}
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1]; data[kern] = []; continue
    if r and r[0] == "Address": hdr = r; continue
    if kern and hdr and len(r) >= len(hdr) - 2:
        data[kern].append(r)
for kern, rs in data.items():
    if want not in kern: continue
    ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
    tot = sum(float(r[ie]) for r in rs); tots = sum(float(r[sm]) for r in rs)
    print("==", kern[:70], "total inst %.0f samples %.0f" % (tot, tots))
    # regions: consecutive instructions whose exec count is within 2x of each other
    regions = []; cur = None
    for i, r in enumerate(rs):
        c = float(r[ie])
        if cur and c > 0 and 0.5 <= c / max(cur["c"], 1) <= 2.0:
            cur["n"] += 1; cur["inst"] += c; cur["samp"] += float(r[sm]); cur["end"] = i
        else:
            if cur: regions.append(cur)
            cur = {"start": i, "end": i, "n": 1, "c": c, "inst": c, "samp": float(r[sm])}
    regions.append(cur)
    for g in sorted(regions, key=lambda g: -g["inst"])[:int(sys.argv[3]) if len(sys.argv) > 3 else 8]:
        print("  [%5d..%5d] n=%4d execs/instr=%.3g inst=%.1f%% samples=%.1f%%  first: %s" % (
            g["start"], g["end"], g["n"], g["c"], 100 * g["inst"] / tot, 100 * g["samp"] / max(tots, 1), rs[g["start"]][1].strip()[:50]))
    if len(sys.argv) > 5:
        a, b = int(sys.argv[4]), int(sys.argv[5])
        for r in rs[a:b]: print("     %8s %6s  %s" % (r[ie], r[sm], r[1].strip()))
