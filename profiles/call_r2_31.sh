#!/bin/bash
# round 2, call 31: side-stream weight gradients at full launch size (experiment)
for V in 0 16; do
CRANK_B200_OPT_ENABLE=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('opt_enable=$V bench', round(d['ms_per_step'],2), round(d['value']), d['gpu_launches']//20)"
done
