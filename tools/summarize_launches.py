"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and by family.
usage: python tools/summarize_launches.py launches.csv [header line ...] > profiles/rNN_ncu_launch_list.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("clb::", "").replace("(int)", "")
    rows.append((name.strip(), us))
tot = sum(u for _, u in rows)
fam = defaultdict(float)
ker = defaultdict(lambda: [0.0, 0])
for n, u in rows:
    base = re.sub(r"<.*$", "", n)
    f = ("attention" if base.startswith("attn_") else "groupnorm" if base.startswith("gn_") else base)
    fam[f] += u
    ker[n][0] += u
    ker[n][1] += 1
for h in sys.argv[2:]:
    print(h)
print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time; per-launch times are cold-cache and serialised: shares are meaningful, absolutes are not\n")
print("-- by family")
for f, u in sorted(fam.items(), key=lambda kv: -kv[1])[:24]:
    print(f"{u:10.1f} us  {100 * u / tot:5.1f} %  {f}")
print("\n-- by kernel")
for n, (u, c) in sorted(ker.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{u:10.1f} us  {100 * u / tot:5.1f} %  x{c:4d}  avg {u / c:8.1f} us  {n}")
