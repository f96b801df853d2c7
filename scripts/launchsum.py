"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = None; agg = {}
for r in rows:
    if len(r) > 5 and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            k = d['Kernel Name'].split('(')[0].replace('void ', '').replace('b200::', '')[:40]
            agg.setdefault(k, []).append(float(d['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if k.startswith('at::'): continue
    print("%-42s n=%3d avg=%9.1f us  share=%.3f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
