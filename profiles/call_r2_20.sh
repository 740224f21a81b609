#!/bin/bash
# round 2, call 20 (N GPUs, N = $NG): weak (64 utts/GPU) and strong (global batch 64) scaling points, eager and graph
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
P=29600
for V in "weak" "strong --global-batch 64" "stronggraph --global-batch 64 --graph"; do
  set -- $V; name=$1; shift
  P=$((P+1))
  timeout 400 $TR --master-port $P bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline "$@" > gpurun_out/r2_scale_${name}_n$NG.json 2> gpurun_out/r2_scale_${name}_n$NG.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_scale_${name}_n$NG.json").read().strip().splitlines()[-1])
    print("$name n=$NG", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["config"]["batch_per_gpu"], "utts/GPU")
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_scale_${name}_n$NG.err").read()[-1200:])
PY
done
