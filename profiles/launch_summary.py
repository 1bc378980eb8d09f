"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time, count and share per kernel.
Usage: python profiles/launch_summary.py <launches.csv> ["command line the list was taken from"]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r[12] != "gpu__time_duration.sum":
        continue
    v = float(r[14].replace(",", ""))
    v = v / 1e6 if r[13] == "ns" else (v / 1e3 if r[13] == "us" else v)      # -> ms
    tot[r[4]] += v
    cnt[r[4]] += 1
T = sum(tot.values())
print((sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]) + f"  ({sum(cnt.values())} launches; serialised, cold cache: compare shares)")
for k, v in tot.most_common(16):
    print(f"{v:10.3f} ms {cnt[k]:5d}x {100 * v / T:5.1f}%  {k[:100]}")
