#!/bin/bash
timeout 600 python bench.py --steps 5 --warmup 3 --no-eager --no-variants --no-frame-path --cpu-budget 0 2>/dev/null | python -c "import json,sys;d=json.load(sys.stdin);print('bs1', d['latency_bs1']['device_p50_ms'], d['latency_bs1']['e2e_host_p50_ms'], d['value'])"
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
