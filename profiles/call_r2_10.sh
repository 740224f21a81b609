#!/bin/bash
# round 2, call 10: round-to-nearest hi/lo split: whole GPU suite + precision diagnostics + bench
mkdir -p gpurun_out
timeout 300 python profiles/diag_r2c.py 2>&1 | grep -v Warn > gpurun_out/r2_diagc_rn.txt; cat gpurun_out/r2_diagc_rn.txt | head -14
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|rel err|differ|worst|log-mel|mlfb|fraction|median" | tail -70 > gpurun_out/r2_pytest_all.log; cat gpurun_out/r2_pytest_all.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_rn.json 2> gpurun_out/r2_bench_rn.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_rn.json"))
    print("bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_rn.err").read()[-1500:])
PY
