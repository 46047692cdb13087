"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log: one line per launch."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < 15: continue
    d.setdefault((r[0], r[4].split('(')[0]), {})[r[12]] = float(r[14])
tot = t = 0
for k, v in d.items():
    rd, wr, ns = v.get('dram__bytes_read.sum', 0), v.get('dram__bytes_write.sum', 0), v.get('gpu__time_duration.sum', 0)
    print(f"{k[1]:42s} {ns/1e6:8.3f} ms  read {rd/1e6:9.1f} MB  write {wr/1e6:9.1f} MB")
    tot += rd + wr; t += ns
print(f"total {t/1e6:.3f} ms, DRAM {tot/1e9:.2f} GB")
