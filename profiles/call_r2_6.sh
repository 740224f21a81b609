#!/bin/bash
# round 2, call 6: graph bit-identity, pre-Adam gradient parity (EMA-warmed), opt-in persistent kernels, conv_pt stamps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_graph.py tests/test_gpu_baseline_shapes.py::test_generator_gradients_match_cpu_oracle_before_adam tests/test_gpu_tc.py::test_persistent_pipelined_kernels_opt_in_path tests/test_gpu_trainstep.py::test_no_grad_forward_and_skipped_final_decoder_are_invisible -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|worst|tensors|assert" | tail -40 > gpurun_out/r2_pytest_6.log; cat gpurun_out/r2_pytest_6.log
CRANK_B200_OPT_ENABLE=3 timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v diag | head -4 > gpurun_out/r2_phases_pt.txt; cat gpurun_out/r2_phases_pt.txt
