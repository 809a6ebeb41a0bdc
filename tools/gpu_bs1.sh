#!/bin/bash
for v in 1 0 1 0; do echo -n "MMT_HALF_SM_SMALL=$v: "; MMT_HALF_SM_SMALL=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-eager --no-variants --no-frame-path --cpu-budget 0 2>/dev/null | python -c "import json,sys;d=json.load(sys.stdin);print('bs1', d['latency_bs1'])"; done
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "graph or pdl or bf16_mode" 2>&1 | tail -2
