"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name and share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = collections.Counter()
cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")[:70]
    per[name] += float(r[-1])
    cnt[name] += 1
tot = sum(per.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us in total")
for k, v in per.most_common():
    print(f"{100 * v / tot:6.2f}%  {v / 1e3:10.1f} us  x{cnt[k]:<4d} {k}")
