#!/bin/bash
# round 2, call 2: new parity tests (BASELINE shapes vs CPU oracle, use_raw, remove_weight_norm, mel tolerance) + graph replay timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_raw_and_wn.py "tests/test_gpu_kernels.py::test_logmel_matches_oracle_on_reference_wav_fixture" "tests/test_gpu_kernels.py::test_offline_mlfb_extraction_matches_reference_fixture" -m gpu -q -s -x 2>&1 | tail -60 > gpurun_out/r2_pytest_new.log
tail -40 gpurun_out/r2_pytest_new.log
for B in 8 64; do
  timeout 300 python bench.py --steps 20 --warmup 3 --batch $B --no-cpu-baseline --no-eager-gpu-baseline --graph > gpurun_out/r2_graph_b$B.json 2> gpurun_out/r2_graph_b$B.err
done
python - <<PY
import json
for n in ("graph_b8", "graph_b64"):
    try:
        d = json.load(open(f"gpurun_out/r2_{n}.json"))
        print(n, round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s e2e", round(d["e2e"]["value"]))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/r2_{n}.err").read()[-1500:])
PY
