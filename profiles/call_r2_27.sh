#!/bin/bash
# round 2, call 27: VQ argmin with 16 warps: parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_shapes.py tests/test_gpu_trainstep.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_27.json 2> gpurun_out/r2_bench_27.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_27.json"))
print("bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()}, d["vq_argmin"]["frac"])
PY
