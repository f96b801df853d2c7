"""Hottest CUDA source lines of an ncu report: joins the SASS source page (pc samples per
instruction) with nvdisasm -g line info of the library's cubin (same build required).
usage: ncu_lines.py report.ncu-rep [top_n] [kernel-substring]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "fqzcomp5_b200", "libb200rans.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
linemap = {}   # mangled -> list of (offset, file, line)
for cub in os.listdir(tmp):
    if "sm_100a" not in cub: continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    fn, cur = None, ("?", 0)
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m: fn = m.group(1); linemap[fn] = {}; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m and fn: linemap[fn][int(m.group(1), 16)] = cur
def demangle(n):
    return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
dm = {demangle(k): k for k in linemap}
src = {}
def srcline(f, ln):
    p = os.path.join(ROOT, "fqzcomp5_b200", "csrc", f)
    if p not in src:
        try: src[p] = open(p).read().splitlines()
        except Exception: src[p] = []
    return src[p][ln - 1].strip() if 0 < ln <= len(src[p]) else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernel, hdr, base, agg, lm = None, None, None, None, None
def flush():
    if agg and want in (kernel or ""):
        tot = sum(v[0] for v in agg.values()) or 1
        toti = sum(v[1] for v in agg.values()) or 1
        print("=== %s  (%d samples, %d warp-instructions)" % (kernel[:80], tot, toti))
        for (f, ln), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print("  %5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * s / tot, 100 * i / toti, f, ln, srcline(f, ln)[:110]))
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        flush(); kernel = r[1]; agg = collections.defaultdict(lambda: [0, 0]); base = None
        def base_name(x):       # "void b200::enc_kernel<(bool)1>(b200::EncJob *, ...)" -> "b200::enc_kernel<1>"
            x = x.replace("(bool)", "").replace(" ", "")
            m2 = re.match(r"(?:void)?([\w:]+(?:<[^()]*>)?)", x)
            return m2.group(1) if m2 else x
        key = [k for d, k in dm.items() if base_name(d) == base_name(kernel)]
        lm = linemap[key[0]] if key else {}
        continue
    if r[0] == "Address": hdr = r; continue
    if hdr and len(r) >= len(hdr) - 1 and r[0].startswith("0x"):
        a = int(r[0], 16)
        if base is None: base = a
        s = float(r[hdr.index("# Samples")] or 0); i = float(r[hdr.index("Instructions Executed")] or 0)
        k = lm.get(a - base, ("?", 0))
        agg[k][0] += s; agg[k][1] += i
flush()
