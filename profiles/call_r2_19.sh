#!/bin/bash
# round 2, call 19 (ONE GPU): BASELINE configs 2 / 4 / 5 (sweep subset), precision modes, graph on/off; compute-sanitizer
mkdir -p gpurun_out
OUT=gpurun_out/r2_sweep.jsonl; : > $OUT
run() { timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline "$@" 2>/dev/null | tail -1 >> $OUT; }
run --trainer vqvae --batch 16                       # config 2
run --trainer cyclegan --batch 64                    # config 4 at 64 utts/GPU
run --trainer cyclegan --batch 8 --graph             # config 4 per-GPU share at 8 GPUs... (cyclegan is not graphed: falls back)
run --batch 8                                        # config 3 per-GPU share at 8 GPUs, eager
run --batch 8 --graph
run --batch 64 --graph
run --precision tf32
run --precision fp32
for T in 128 512 2048; do for B in 8 64 256; do
  if [ $((B*T)) -le 140000 ]; then run --batch $B --frames $T; fi
done; done
run --batch 8 --frames 4096
python - <<PY
import json
for l in open("$OUT"):
    try:
        d = json.loads(l)
        c = d["config"]
        print(c["trainer"], "B", c["batch_per_gpu"], "T", c["frames"], d["dtype"], "graph" if c["cuda_graph"] else "eager", round(d["ms_per_step"], 2), "ms", round(d["value"]), "frames/s")
    except Exception as e:
        print("bad line", e, l[:100])
PY
# sanitizer: memcheck over the kernel / tensor-core / train-step / data / eval tests at their (small) test shapes,
# racecheck over the kernel tests
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_trainstep.py tests/test_gpu_data.py tests/test_gpu_eval.py tests/test_gpu_raw_and_wn.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2_sanitizer_memcheck.log; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log; tail -6 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py -m gpu -q -x -k "not baseline" 2>&1 | tail -12 > gpurun_out/r2_sanitizer_racecheck.log; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log; tail -6 gpurun_out/r2_sanitizer_racecheck.log
