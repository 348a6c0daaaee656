#!/bin/bash
# full BASELINE configs[2] (25 MS/s, 36 ch, 300 s = 30 GB per step) and the configs[3] slice, with the oracle parity check
tag=$1
timeout 900 python bench.py --no-cpu-baseline --workload cfg3 > gpurun_out/${tag}_bench_cfg3_full.json 2> gpurun_out/${tag}_bench_cfg3_full.err
timeout 400 python bench.py --no-cpu-baseline --workload cfg4s > gpurun_out/${tag}_bench_cfg4s.json 2> gpurun_out/${tag}_bench_cfg4s.err
python - <<PY
import json
for w in ("cfg3_full", "cfg4s"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % w)); r = d["roofline"]
        print(w, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "plan", round(r["planner_ms_per_step"], 2), "synth", round(r["synth_ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "parity", d["parity_check"]["differing_samples"], "of blocks", d["parity_check"]["blocks"], r["kernel"])
    except Exception as e:
        print(w, "failed", e)
PY
