"""One-screen summary of a bench.py JSON line.  usage: benchsum.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))


def show(name, e):
    print("%-18s value %6.1f  enc %6.1f dec %6.1f (packed enc %s)  e2e %5.1f (enc %.1f dec %.1f)  step %.2f ms" % (
        name, e["value"], e["enc_gbs"], e["dec_gbs"], "%.1f" % e["enc_packed_gbs"] if e.get("enc_packed_gbs") else "-",
        e["e2e"]["value"], e["e2e"]["enc_gbs"], e["e2e"]["dec_gbs"], e["ms_per_step"]))
    rs = e["roofline_step"]
    print("   kernels: enc %.3f ms (%.4f)  dec %.3f ms (%.4f) | steps: enc %.3f ms (%.4f)  dec %.3f ms (%.4f)" % (
        e["roofline_enc"]["kernel_ms"], e["roofline_enc"]["frac"], e["roofline_dec"]["kernel_ms"],
        e["roofline_dec"]["frac"], rs["enc_step_ms"], rs["enc_step_frac"], rs["dec_step_ms"], rs["dec_step_frac"]))
    cb = e.get("cpu_baseline", {})
    print("   cpu %s  scalar %s  parity slices %s  flags %s  launches %s  clocks %s" % (
        "%.2f" % cb["value"] if cb else "-", "%.2f" % cb["as_shipped_scalar"]["value"] if cb.get("as_shipped_scalar", {}).get("value") else "-",
        e.get("parity_checked_slices"), e.get("stream_flags"), e.get("gpu_launches"), e["clocks"]))


if "roofline_step" in d:
    show("headline", d)
for k, v in d.get("per_config", {}).items():
    if "roofline_step" in v:
        show(k, v)
b = d.get("per_config", {}).get("fastq_blocks_m3") or (d if "variants" in d else None)
if b:
    for vn, v in b["variants"].items():
        if "picked" in v:
            print("blocks %-20s value %.2f enc %.2f dec %.2f GB/s text | picked %s" % (vn, v["value"], v["enc_gbs"], v["dec_gbs"], v["picked"]))
            continue
        print("blocks %-12s value %.2f enc %.2f dec %.2f GB/s text | blocks %d x %.2f GB -> ratio %.3f | wins %s | phases %s" % (
            vn, v["value"], v["enc_gbs"], v["dec_gbs"], v["blocks"], v["block_bytes"] / 1e9, v["ratio"], v["wins"],
            {k: round(x, 1) for k, x in v.get("encode_phase_ms_per_block", {}).items()}))
    cb = b.get("cpu_baseline")
    if cb:
        print("   cpu codec-only", {k: round(v["value"], 2) for k, v in cb["codec_only"].items()},
              "tool", {k: (round(v["enc_gbs"], 3), round(v["dec_gbs"], 3)) for k, v in cb["e2e_tool"].items() if isinstance(v, dict) and "enc_gbs" in v})
