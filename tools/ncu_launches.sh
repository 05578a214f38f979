#!/bin/bash
# Per-launch device time of one short bench run (cold-cache, serialised: compare SHARES, not absolutes).
# Usage (under gpurun): bash tools/ncu_launches.sh <tag>
set -e
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches_${tag}.log 2>&1 || true
python - <<PY
import csv, collections, json
rows = [r for r in csv.reader(open("gpurun_out/launches_${tag}.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
out = [{"kernel": k, "launches": n, "total_us": round(v / 1e3, 1) if v > 1e5 else round(v, 1), "share": round(v / tot, 4)} for k, (n, v) in agg.items()]
json.dump(out, open("gpurun_out/launches_${tag}.json", "w"), indent=1)
for o in sorted(out, key=lambda o: -o["share"]):
    print(o)
PY
