#!/bin/bash
# round 2, call 23: weight gradients of small launches on a side stream: GPU suite + 8-utterance bench (eager / graph, on / off) + 64
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_pytest_23.log; cat gpurun_out/r2_pytest_23.log
run() { name=$1; shift; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline "$@" 2>/dev/null > gpurun_out/r2_bench_23_$name.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_23_$name.json')); print('$name', round(d['ms_per_step'],2), 'ms', round(d['value']), 'frames/s')"; }
run b8_eager --batch 8
run b8_graph --batch 8 --graph
CRANK_B200_OPT_DISABLE=512 run b8_eager_noside --batch 8
CRANK_B200_OPT_DISABLE=512 run b8_graph_noside --batch 8 --graph
run b16_vqvae --batch 16 --trainer vqvae
run b64 --batch 64
