#!/bin/bash
# round 2, call 12: log-mel backward / learnable windows / Griffin-Lim eval tests, re-timed fused log-mel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_eval.py -m gpu -q -s -k "logmel or mlfb or learnable or griffin or wavs" 2>&1 | grep -E "passed|failed|FAILED|Error|error|rel|log-mel|mlfb|assert" | tail -40 > gpurun_out/r2_pytest_logmel2.log; cat gpurun_out/r2_pytest_logmel2.log
timeout 300 python bench.py --workload logmel --steps 20 --warmup 3 > gpurun_out/r2_bench_logmel.json 2> gpurun_out/r2_bench_logmel.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_logmel.json')); print('logmel', d['ms_per_step'], d['roofline']['avg_us'], d['roofline']['frac'], d['cufft_path']['ms_per_step'])"; tail -3 gpurun_out/r2_bench_logmel.err
