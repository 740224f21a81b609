#!/bin/bash
# round 2, call 32 (2 GPUs): data parallelism with the final kernels (side-stream weight gradients, batched D pass): equivalence + one weak point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29711 profiles/dp_equiv.py 2>&1 | tail -3
cp gpurun_out/dp_equiv_lsgan_2gpu.json gpurun_out/r2_dp_equiv_final_2gpu.json 2>/dev/null
timeout 300 $TR --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_scale_weakfinal_n2.json 2> gpurun_out/r2_scale_weakfinal_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_scale_weakfinal_n2.json').read().strip().splitlines()[-1]); print('weak n=2 final build', round(d['ms_per_step'],2), round(d['value']))" || tail -20 gpurun_out/r2_scale_weakfinal_n2.err
