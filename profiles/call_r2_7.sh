#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/diag_r2b.py fp32 16 500 1 2>&1 | grep -v Warn | tail -60 > gpurun_out/r2_diagb_fp32.txt; cat gpurun_out/r2_diagb_fp32.txt
timeout 300 python profiles/diag_r2b.py tf32x3 16 500 1 2>&1 | grep -v Warn | tail -60 > gpurun_out/r2_diagb_tf32x3.txt; cat gpurun_out/r2_diagb_tf32x3.txt
timeout 300 python profiles/diag_r2b.py fp32 2 96 1 2>&1 | grep -v Warn | tail -60 > gpurun_out/r2_diagb_fp32_toy.txt; cat gpurun_out/r2_diagb_fp32_toy.txt
