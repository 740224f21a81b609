#!/bin/bash
mkdir -p gpurun_out
timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 profiles/dp_equiv.py 2>&1 | grep -E "^\{" | tail -1
cp gpurun_out/dp_equiv_lsgan_2gpu.json gpurun_out/r2_dp_equiv_final2_2gpu.json 2>/dev/null
