#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/diag_r2c.py 2>&1 | grep -v Warn > gpurun_out/r2_diagc.txt; cat gpurun_out/r2_diagc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dp_equiv.py 2>&1 | tail -5
