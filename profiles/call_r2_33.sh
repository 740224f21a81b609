#!/bin/bash
mkdir -p gpurun_out
timeout 100 python profiles/side_stream_check.py 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/r2_side_stream_check.txt
