#!/bin/bash
# round 2, call 15 (2 GPUs): DP equivalence with the real kernels, weak + strong scaling points at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 profiles/dp_equiv.py 2>&1 | tail -4
cp gpurun_out/dp_equiv_lsgan_2gpu.json gpurun_out/r2_dp_equiv_lsgan_2gpu.json 2>/dev/null
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_scale_weak_n2.json 2> gpurun_out/r2_scale_weak_n2.err
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --global-batch 64 > gpurun_out/r2_scale_strong_n2.json 2> gpurun_out/r2_scale_strong_n2.err
timeout 400 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --global-batch 64 --graph > gpurun_out/r2_scale_strong_graph_n2.json 2> gpurun_out/r2_scale_strong_graph_n2.err
python - <<PY
import json
for n in ("weak_n2", "strong_n2", "strong_graph_n2"):
    try:
        d = json.load(open(f"gpurun_out/r2_scale_{n}.json"))
        print(n, round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d["config"]["batch_per_gpu"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/r2_scale_{n}.err").read()[-800:])
PY
