#!/bin/bash
# One GPU round trip: parity tests, bench cfg2 + cfg3s, optional ncu capture.  Usage (under gpurun):
#   tools/gpu_check.sh <tag> [ncu]
tag=$1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
timeout 300 python bench.py --no-cpu-baseline --workload cfg3s > gpurun_out/${tag}_bench_cfg3s.json 2> gpurun_out/${tag}_bench_cfg3s.err
python - <<PY
import json
for w in ("cfg2", "cfg3s"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % w)); r = d["roofline"]
        print(w, "value", round(d["value"]), "plan", round(r["planner_ms_per_step"], 2), "synth", round(r["synth_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "fallback", d.get("exact_fallback_samples"))
    except Exception as e:
        print(w, "failed", e)
PY
if [ "$2" = "ncu" ]; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:e1_synth -c 1 -o gpurun_out/${tag}_synth_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
    tail -1 gpurun_out/${tag}_ncu.log
fi
