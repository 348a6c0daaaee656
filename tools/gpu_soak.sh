#!/bin/bash
# sanitizer runs of both synthesis kernels + the soak (bench-size streams repeated, compared bit for bit with the oracle)
tag=$1; reps=${2:-3}
{
  echo "# compute-sanitizer on tools/sanitize_case.py (B200, ${tag} build): carry-walked kernel, then the event-driven one"
  for k in "" ev; do
    for tool in memcheck racecheck; do
      timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py $k 2>&1 | grep -E "equal|SUMMARY|ERROR|hazard" | head -8
    done
  done
} > gpurun_out/${tag}_sanitizer.txt 2>&1
cat gpurun_out/${tag}_sanitizer.txt
timeout 1500 python tools/find_mismatch.py $reps > gpurun_out/${tag}_soak.jsonl 2> gpurun_out/${tag}_soak.err
echo "soak exit $?"
python - <<PY
import json
tot = 0
for l in open("gpurun_out/${tag}_soak.jsonl"):
    d = json.loads(l); tot += d["samples"] * d["reps"]
    print(d["case"], d["samples"], d["reps"], d["differing_per_rep"])
print("samples compared", tot)
PY
