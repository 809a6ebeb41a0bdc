#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
timeout 600 python tools/bench_gemm_ln.py 2>&1 | tail -6 | tee gpurun_out/gemm_exp1.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; python -c "import json;d=json.load(open('gpurun_out/r2x_bench.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'bs1', d['latency_bs1']['device_p50_ms'], d['latency_bs1']['e2e_host_p50_ms'], 'e2e', round(d['e2e']['value'],1), d['clocks'], d['roofline'])"
