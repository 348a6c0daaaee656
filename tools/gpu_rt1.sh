#!/bin/bash
# real-time call shape: parity tests, rt1 bench (coarse spans vs per-pass geometry), launch list
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests.log
E1B200_COARSE_SPANS=1 timeout 300 python bench.py --no-cpu-baseline --workload rt1 --steps 200 --warmup 10 > gpurun_out/${tag}_bench_rt1_coarse.json 2> gpurun_out/${tag}_bench_rt1_coarse.err
timeout 300 python bench.py --no-cpu-baseline --workload rt1 --steps 200 --warmup 10 > gpurun_out/${tag}_bench_rt1.json 2> gpurun_out/${tag}_bench_rt1.err
python - <<PY
import json
for w in ("rt1_coarse", "rt1"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % w)); r = d["roofline"]
        print(w, "ms/call device", round(d["ms_per_step"], 3), "plan", round(r["planner_ms_per_step"], 3), "synth", round(r["synth_ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d["parity_check"]["differing_samples"])
    except Exception as e:
        print(w, "failed", e)
PY
bash tools/gpu_launches.sh ${tag} rt1 | tail -14
