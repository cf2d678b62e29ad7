"""Summarise an ncu --csv launch list: python tools/launch_summary.py launches.csv [filter-substring ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cur = {}
for r in rows[hi + 1:]:
    cur.setdefault((int(r[0]), r[4].split("(")[0][-44:]), {})[r[12]] = r[14]
filt = sys.argv[2:]
for (i, name), m in sorted(cur.items()):
    if not filt or any(f in name for f in filt):
        t = float(m.get("gpu__time_duration.sum", 0)) / 1e6
        rd = float(m.get("dram__bytes_read.sum", 0)) / 1e9
        wr = float(m.get("dram__bytes_write.sum", 0)) / 1e9
        print(f"{i:4d} {name:44s} {t:9.3f} ms  read {rd:8.2f} GB  write {wr:8.2f} GB  {(rd + wr) / max(t, 1e-9):8.1f} GB/s*1e3")
