"""Per-launch table from an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` log."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:48]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in sorted(d.items())[:int(sys.argv[2]) if len(sys.argv) > 2 else 60]:
    print("%3d %-48s %9.1f us  R %6.0f MB  W %6.0f MB" % (i, k, m['gpu__time_duration.sum'] / 1e3,
          m.get('dram__bytes_read.sum', 0) / 1e6, m.get('dram__bytes_write.sum', 0) / 1e6))
