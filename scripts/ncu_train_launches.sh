#!/bin/bash
# ncu launch list of the train step (scripts/bench_train.py): per-kernel device time of the last third of the launches
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv python scripts/bench_train.py --steps 4 --warmup 8 > gpurun_out/ncu_train.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/train_launches.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = rows[hi + 1:]
agg = collections.OrderedDict()
n = len(data)
for r in data[n - n // 3:]:
    v = float(r[vi].replace(",", "")); v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(r[ui], v)
    k = r[ki].split("(")[0].replace("void ", "")[:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
out = ["%-60s %4d %10.1f us total %8.1f avg  %.3f" % (k, v[0], v[1], v[1] / v[0], v[1] / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
out.append("launches considered %d, total %.1f us" % (n // 3, tot))
print("\n".join(out))
open("gpurun_out/train_launch_list.txt", "w").write("\n".join(out) + "\n")
PY
