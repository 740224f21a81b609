#!/bin/bash
mkdir -p gpurun_out
timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v Warn | grep -B2 "wgrad" > gpurun_out/r2_phases_wgrad.txt; cat gpurun_out/r2_phases_wgrad.txt
