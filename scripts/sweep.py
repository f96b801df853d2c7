"""Stream-decomposition sweep: K calls of S bytes for each workload.
Writes profiles/r1_sweep.jsonl and profiles/r1_sweep.md."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
slices = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20]
for wl in ["illumina_qual_o0", "illumina_qual_o1", "ont_qual_o1", "illumina_seq_c5"]:
    for S in slices:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--slice", str(S),
               "--steps", "3", "--warmup", "3", "--no-cpu"]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900).stdout
        for l in out.splitlines():
            if l.startswith("{"):
                d = json.loads(l)
                rows.append(d)
                print(wl, S, "%.1f GB/s" % d["value"], flush=True)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "r1_sweep.jsonl"), "w") as f:
    for d in rows:
        f.write(json.dumps(d) + "\n")
with open(os.path.join(ROOT, "profiles", "r1_sweep.md"), "w") as f:
    f.write("# Stream-decomposition sweep, 1 GB block, one B200 (device-resident unless noted), GB/s of uncompressed data\n\n")
    f.write("| workload | S (bytes per call) | K (calls) | C/U | enc | dec | round trip | e2e round trip | enc kernel ms | dec kernel ms |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for d in rows:
        c = d["config"]
        f.write("| %s | %d | %d | %.3f | %.0f | %.0f | %.0f | %.1f | %.2f | %.2f |\n" % (
            c["workload"], c["slice_bytes"], c["streams"], c["ratio"], d["enc_gbs"], d["dec_gbs"], d["value"],
            d["e2e"]["value"], d["roofline_enc"]["kernel_ms"], d["roofline_dec"]["kernel_ms"]))
print("done")
