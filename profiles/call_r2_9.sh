#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/diag_r2d.py 16 500 2>&1 | grep -v Warn > gpurun_out/r2_diagd.txt; cat gpurun_out/r2_diagd.txt
timeout 300 python profiles/diag_r2d.py 2 150 2>&1 | grep -v Warn > gpurun_out/r2_diagd_small.txt; cat gpurun_out/r2_diagd_small.txt
