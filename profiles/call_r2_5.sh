#!/bin/bash
# round 2, call 5: swizzled MN-major probe, GPU suite after the save_gates fix, persistent kernels with prefetched epilogues
mkdir -p gpurun_out
timeout 60 profiles/hwprobe/hwprobe T7 2>&1 | grep -v "^== T3" > gpurun_out/hwprobe_r2_t7.txt; cat gpurun_out/hwprobe_r2_t7.txt
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_trainstep.py tests/test_gpu_baseline_shapes.py tests/test_gpu_raw_and_wn.py tests/test_gpu_kernels.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|rel err|differ|worst|log-mel|mlfb|fraction" | tail -60 > gpurun_out/r2_pytest_all.log; cat gpurun_out/r2_pytest_all.log
timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v diag | head -3 > gpurun_out/r2_phases_pt.txt; cat gpurun_out/r2_phases_pt.txt
for M in 0 16 32 48; do
  CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_mask$M.json 2> gpurun_out/r2_bench_mask$M.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_mask$M.json"))
    print("mask $M", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("mask $M failed", e); print(open("gpurun_out/r2_bench_mask$M.err").read()[-1500:])
PY
done
