#!/bin/bash
# round 2, call 24 (ONE GPU): round-end evidence: GPU suite, smoke, default bench line (CPU + eager-GPU arms), log-mel line, phase stamps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_final.log; cat gpurun_out/r2_pytest_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 400 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null
timeout 300 python bench.py --workload logmel --steps 20 --warmup 3 > gpurun_out/r2_bench_logmel.json 2>/dev/null
timeout 200 python profiles/phase_probe.py tf32x3 2>&1 | grep -v Warn > gpurun_out/r2_phases_final.txt; cat gpurun_out/r2_phases_final.txt
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_final.json"))
print("final", round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "steps", d["steps"], "warmup", d["warmup"])
print(" kernels", {k: (round(v["ms_per_step"], 2), round(v["tflops"] or 0, 1)) for k, v in d["kernels"].items()})
print(" roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["all_dense_kernels"])
print(" vq", d["vq_argmin"]["frac"], " cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"), " eager", (d.get("torch_eager_gpu_baseline") or {}).get("value"), (d.get("torch_eager_gpu_baseline") or {}).get("speedup_of_this_repo"))
r = json.load(open("gpurun_out/r2_bench_reference.json")); print(" reference arm", round(r["value"]), r["cpu_baseline"]["cores"], r["config"]["batch_per_gpu"])
l = json.load(open("gpurun_out/r2_bench_logmel.json")); print(" logmel", round(l["value"]), l["roofline"]["avg_us"], l["roofline"]["frac"])
PY
