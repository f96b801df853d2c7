"""Pretty-print the interesting parts of bench.py's JSON line (stdin)."""
import json, sys
for l in sys.stdin:
    if not l.startswith("{"):
        print(l, end="")
        continue
    d = json.loads(l)
    print("value %.1f %s  ms/step %.3f  enc %.1f  dec %.1f  launches %s" % (
        d["value"], d["unit"], d["ms_per_step"], d.get("enc_gbs", 0), d.get("dec_gbs", 0), d.get("gpu_launches")))
    if "e2e" in d:
        e = d["e2e"]; print("e2e %.2f (enc %.2f dec %.2f)" % (e["value"], e.get("enc_gbs", 0), e.get("dec_gbs", 0)))
    for k in ("roofline_enc", "roofline_dec"):
        if k in d: print(k, "kernel_ms %.3f achieved %.0f GB/s frac %.3f" % (d[k]["kernel_ms"], d[k]["achieved"], d[k]["frac"]))
    print("config", d["config"]); print("clocks", d.get("clocks"))
    if "cpu_baseline" in d: print("cpu", d["cpu_baseline"])
