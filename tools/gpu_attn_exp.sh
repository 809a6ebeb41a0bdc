#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -5
for v in "" half; do for mode in sym cross; do echo -n "'$v' $mode: "; MMT_B200_DEV_LIB=$v timeout 300 python tools/bench_attn.py $mode 2>&1 | tail -1; done; done 2>&1 | tee gpurun_out/attn_exp5.txt
