#!/bin/bash
# quick perf iteration: phase probe + one short bench (no pytest)
mkdir -p gpurun_out
timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v diag > gpurun_out/phases_tf32x3.txt; cat gpurun_out/phases_tf32x3.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print(round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
PY
