#!/bin/bash
# round 2, call 16: 2-CTA/SM fused forward (k_resblock_fwd_tc2): parity suite + bench A/B (opt-disable 64 = round-1 kernel)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_baseline_shapes.py tests/test_gpu_trainstep.py tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest_16.log; cat gpurun_out/r2_pytest_16.log
for M in 0 64; do
CRANK_B200_OPT_DISABLE=$M timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_bench_16_m$M.json 2> gpurun_out/r2_bench_16_m$M.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_16_m$M.json"))
    print("mask $M bench", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["gpu_launches"], {k: (round(v["ms_per_step"], 2), round(v["avg_us"],1)) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_16_m$M.err").read()[-1500:])
PY
done
timeout 200 python bench.py --steps 10 --warmup 3 --batch 8 --graph --no-cpu-baseline --no-eager-gpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b8 graph', round(d['ms_per_step'],2))"
