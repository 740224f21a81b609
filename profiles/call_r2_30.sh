#!/bin/bash
# round 2, call 30: one discriminator pass over [real | fake] in the D update: parity + A/B
timeout 1200 python -m pytest tests/test_gpu_trainstep.py tests/test_gpu_tc.py tests/test_gpu_baseline_shapes.py tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -3
for V in 1 0; do
CRANK_B200_BATCH_D=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch_D=$V bench', round(d['ms_per_step'],2), round(d['value']), d['gpu_launches']//20, {k: round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"
done
