#!/bin/bash
# ncu launch list (per-kernel durations, cold-cache and serialised) of one bench step.  Usage: tools/gpu_launches.sh <tag> <workload>
tag=$1; wl=${2:-cfg2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_${wl}.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${tag}_launch.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(l for l in open("gpurun_out/${tag}_launches_${wl}.csv") if l.startswith('"')))
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[-24:]:
    print(r[ki][:48].ljust(48), r[vi], r[h.index("Grid Size")], r[h.index("Block Size")])
PY
