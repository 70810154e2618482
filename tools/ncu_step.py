"""Summarise one training step out of an ncu launch list.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python bench.py --steps 1 --warmup 3 --two-streams 0 --no-graph --no-cpu-baseline
    python tools/ncu_step.py launches.csv > profiles/<round>_launches_by_kernel.txt

A step is delimited by its two adam_kernel launches (discriminator, then generator update); the last complete step of
the file's first half is reported, or step `argv[2]` (0 for a `bench.py --ncu-step` file, which holds exactly one)
(eager, one stream: ncu serialises launches and flushes caches, so compare shares)."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((r["Kernel Name"], us))
adam = [i for i, (k, _) in enumerate(rows) if "adam_kernel" in k]
pick = int(sys.argv[2]) if len(sys.argv) > 2 else max(0, len(adam) // 2 - 1) // 2 * 2
if len(adam) < pick + 2:
    sys.exit("not enough adam launches to cut a step: %d" % len(adam))
lo, hi = adam[pick - 1] + 1 if pick > 0 else 0, adam[pick + 1] + 1
step = rows[lo:hi]
agg = OrderedDict()
for k, us in step:
    k = re.sub(r"\(.*$", "", k).strip()
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none; one eager training step (launches %d..%d of %d)" % (lo, hi, len(rows)))
print("# %d launches, %.1f ms summed (cold-cache, serialised: compare shares)" % (len(step), tot / 1000.0))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s x%-5d %9.1f us  %5.1f%%" % (k[:70], n, us, 100.0 * us / tot))
