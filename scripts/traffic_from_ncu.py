"""profiles/traffic.json from the ncu CSVs of scripts/traffic_probe.py: DRAM bytes (read + written) and duration of
every launch of one encode step and one decode step, per workload.
usage: traffic_from_ncu.py out.json name=csv [name=csv ...]"""
import csv
import json
import sys

out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of ONE encode step and ONE decode step "
               "(scripts/traffic_probe.py under ncu --clock-control none; 1 GB block, 256 KiB streams); "
               "enc_kernel / dec_kernel = the coder launches b200rans_last_kernel_ms brackets, *_step = every launch "
               "of the step; torch's own kernels (index / compare) are left out"}
for arg in sys.argv[2:]:
    name, path = arg.split("=")
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = None
    launches = {}
    order = []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            k = d["Kernel Name"]
            if k.startswith("void at::") or k.startswith("at::") or "elementwise" in k or "index" in k.lower():
                continue
            kid = d["ID"]
            if kid not in launches:
                launches[kid] = {"name": k.split("(")[0].replace("void ", "").replace("b200::", ""), "bytes": 0.0, "ns": 0.0}
                order.append(kid)
            v = float(d["Metric Value"].replace(",", ""))
            unit = d["Metric Unit"]
            if d["Metric Name"].startswith("dram__bytes"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                launches[kid]["bytes"] += v * mult
            elif d["Metric Name"] == "gpu__time_duration.sum":
                mult = {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}[unit]
                launches[kid]["ns"] += v * mult
    enc_names = ("hist_kernel", "prep_kernel", "enc_kernel", "scan_kernel", "gather_kernel", "inslot_results", "assemble_parents",
                 "stripe", "trial", "rcp_table")
    e = {"enc_kernel": 0, "dec_kernel": 0, "enc_step": 0, "dec_step": 0, "launches": []}
    for kid in order:
        L = launches[kid]
        nm = L["name"]
        is_dec = nm.startswith("dec_")
        e["launches"].append({"kernel": nm, "dram_bytes": int(L["bytes"]), "us": round(L["ns"] / 1e3, 1)})
        if is_dec:
            e["dec_step"] += int(L["bytes"])
            if nm != "dec_results_kernel":
                e["dec_kernel"] += int(L["bytes"])
        elif any(nm.startswith(x) for x in enc_names):
            e["enc_step"] += int(L["bytes"])
            if nm.startswith("enc_kernel"):
                e["enc_kernel"] += int(L["bytes"])
    e["source"] = "profiles/" + path.split("/")[-1]
    out[name] = e
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "launches"} for k, v in out.items() if k != "note"}, indent=1))
