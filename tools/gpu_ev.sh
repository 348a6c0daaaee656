#!/bin/bash
# GPU round trip for the event-driven kernel: parity tests, cfg3s bench (with the oracle parity check), optional ncu capture on cfg3s.
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1
tail -3 gpurun_out/${tag}_gpu_tests.log
timeout 400 python bench.py --no-cpu-baseline --workload cfg3s > gpurun_out/${tag}_bench_cfg3s.json 2> gpurun_out/${tag}_bench_cfg3s.err
tail -2 gpurun_out/${tag}_bench_cfg3s.err
python - <<PY
import json
for w in ("cfg3s",):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % w)); r = d["roofline"]
        print(w, "value", round(d["value"]), "plan", round(r["planner_ms_per_step"], 2), "synth", round(r["synth_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "fallback", d.get("exact_fallback_samples"), "parity", d.get("parity_check"))
    except Exception as e:
        print(w, "failed", e)
PY
if [ "$2" = "ncu" ]; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:e1_synth -c 1 -o gpurun_out/${tag}_synth_cfg3s_full python bench.py --workload cfg3s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${tag}_ncu.log 2>&1
    tail -1 gpurun_out/${tag}_ncu.log
fi
