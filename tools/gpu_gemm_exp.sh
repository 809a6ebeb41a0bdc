#!/bin/bash
mkdir -p gpurun_out
for v in "" st3 "" st3; do echo "== variant '$v'"; MMT_B200_DEV_LIB=$v timeout 600 python tools/bench_gemm_ln.py 2>&1 | tail -4; done | tee gpurun_out/gemm_exp2.txt
