#!/bin/bash
# round 2, call 29: sanity of the final tree: GPU suite + smoke + short bench
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['ms_per_step'],2), round(d['value']), {k: round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"
