#!/bin/bash
# round-end evidence in one gpurun call (ONE GPU): parity suite, default bench (+ CPU baseline), workload variants,
# ncu launch list of one step, one --set full capture per tensor-core kernel family
set -x
mkdir -p gpurun_out
PREC=tf32x3
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
timeout 120 python bench.py --steps 10 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>/dev/null
timeout 120 python bench.py --steps 10 --warmup 3 --trainer cyclegan --no-cpu-baseline > gpurun_out/bench_cyclegan.json 2>/dev/null
timeout 120 python bench.py --steps 10 --warmup 3 --trainer vqvae --batch 16 --no-cpu-baseline > gpurun_out/bench_vqvae16.json 2>/dev/null
timeout 120 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2>/dev/null
python - <<PY
import json
for n in ("default", "b8", "cyclegan", "vqvae16", "tf32"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]), d.get("cpu_baseline", {}).get("value"), (d.get("data_path") or {}).get("ms_per_batch_wall"))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 900 --csv \
    --log-file gpurun_out/launches_${PREC}.csv python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline > gpurun_out/ncu_bench_${PREC}.log 2>&1
for K in k_resblock_fwd_tc k_conv_tc k_wgrad_tc_raw; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:${K}\\b -s 40 -c 1 \
    -o gpurun_out/prof_${K}_${PREC} -f python bench.py --steps 1 --warmup 3 --precision $PREC --no-cpu-baseline > gpurun_out/ncu_full_${K}.log 2>&1
done
ls gpurun_out | head -40
