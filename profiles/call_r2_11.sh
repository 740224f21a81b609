#!/bin/bash
# round 2, call 11: fused log-mel: parity tests, bench line, ncu full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "logmel or mlfb" 2>&1 | grep -E "passed|failed|FAILED|Error|rel|log-mel|mlfb" | tail -30 > gpurun_out/r2_pytest_logmel.log; cat gpurun_out/r2_pytest_logmel.log
timeout 300 python bench.py --workload logmel --steps 20 --warmup 3 > gpurun_out/r2_bench_logmel.json 2> gpurun_out/r2_bench_logmel.err; cat gpurun_out/r2_bench_logmel.json; tail -5 gpurun_out/r2_bench_logmel.err
timeout 300 python bench.py --workload logmel --batch 256 --frames 2000 --steps 20 --warmup 3 > gpurun_out/r2_bench_logmel_big.json 2>> gpurun_out/r2_bench_logmel.err; cat gpurun_out/r2_bench_logmel_big.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_logmel -c 1 -s 3 -f -o gpurun_out/prof_r2_k_logmel python bench.py --workload logmel --steps 3 --warmup 3 > gpurun_out/ncu_logmel.log 2>&1; tail -3 gpurun_out/ncu_logmel.log
