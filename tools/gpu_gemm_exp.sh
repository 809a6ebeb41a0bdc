#!/bin/bash
mkdir -p gpurun_out
MMT_B200_DEV_LIB=rcpa timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
for v in "" rcpa "" rcpa; do echo "== variant '$v'"; MMT_B200_DEV_LIB=$v timeout 600 python tools/bench_gemm_ln.py 2>&1 | tail -2; done | tee gpurun_out/gemm_exp4.txt
for v in "" rcpa; do echo -n "step '$v': "; MMT_B200_DEV_LIB=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 2>/dev/null | python -c "import json,sys;d=json.load(sys.stdin);print(round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; done
