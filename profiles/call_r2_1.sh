#!/bin/bash
# round 2, call 1: hardware probe + whole-step CUDA graph on hardware + eager-GPU bar
mkdir -p gpurun_out
for g in T1v0 T1v1 T1v2 T3a T3u T4 T5 T6; do
  timeout 60 profiles/hwprobe/hwprobe $g 2>&1 | grep -v "^==" ; echo "[group $g exit $?]"
done > gpurun_out/hwprobe_r2.txt 2>&1
cat gpurun_out/hwprobe_r2.txt | tail -150
for B in 8 64; do
  timeout 300 python bench.py --steps 10 --warmup 3 --batch $B --no-cpu-baseline --no-eager-gpu-baseline > gpurun_out/r2_eager_b$B.json 2> gpurun_out/r2_eager_b$B.err
  timeout 300 python bench.py --steps 10 --warmup 3 --batch $B --no-cpu-baseline --no-eager-gpu-baseline --graph > gpurun_out/r2_graph_b$B.json 2> gpurun_out/r2_graph_b$B.err
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_torch_eager_bar.json 2> gpurun_out/r2_torch_eager_bar.err
python - <<PY
import json
for n in ("eager_b8", "graph_b8", "eager_b64", "graph_b64", "torch_eager_bar"):
    try:
        d = json.load(open(f"gpurun_out/r2_{n}.json"))
        print(n, round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "frames/s", d.get("gpu_launches"), d.get("torch_eager_gpu_baseline"))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/r2_{n}.err").read()[-1500:])
PY
