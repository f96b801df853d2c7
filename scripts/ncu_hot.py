"""Summarise an ncu report: per-kernel key metrics + hottest source lines (by pc samples)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(r, name):
    return r[hdr.index(name)] if name in hdr else "?"
for r in rows[2:]:
    print("==", col(r, "Kernel Name")[:60], "dur", col(r, "gpu__time_duration.sum"), "regs", col(r, "launch__registers_per_thread"),
          "warps_active%", col(r, "sm__warps_active.avg.pct_of_peak_sustained_active")[:5],
          "issue%", col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")[:5],
          "inst", col(r, "smsp__inst_executed.sum"), "dramR", col(r, "dram__bytes_read.sum"), "dramW", col(r, "dram__bytes_write.sum"))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and r[i].replace('.','').isdigit()]
    tot = sum(v for v, _ in st) or 1
    print("   stalls:", ", ".join("%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for v, h in sorted(st, reverse=True)[:6]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page prints one table per kernel; aggregate "# Samples" by source line if present
cur = None
for block in src.split("\n\n"):
    lines = block.strip().splitlines()
    if not lines: continue
    rd = list(csv.reader(lines))
    h = rd[0]
    if "Source" not in h: 
        continue
    scol = [i for i, x in enumerate(h) if "Sampl" in x and "Not" not in x]
    if not scol: continue
    si = scol[0]
    agg = []
    for r in rd[1:]:
        try: agg.append((float(r[si]), r[h.index("Source")][:110]))
        except Exception: pass
    tot = sum(v for v, _ in agg) or 1
    print("--- hottest lines (of %d samples)" % tot)
    for v, s_ in sorted(agg, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
        print("   %5.1f%%  %s" % (100 * v / tot, s_.strip()))
