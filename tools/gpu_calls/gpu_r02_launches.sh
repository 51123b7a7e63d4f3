#!/bin/bash
# Round 2: ncu launch list of the headline bench command on the final build (shares of the step per kernel).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_roundend.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02_launches_roundend.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02_launches_roundend.csv") if not l.startswith("==")))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6}.get(u, 1e-6)
    name = r[ki].split("(")[0]
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(8): print(f"{k}: {cnt[k]} launches, {v:.2f} ms, {100 * v / s:.1f} %")
PY
