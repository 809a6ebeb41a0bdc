#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3
for mode in sym cross; do echo -n "$mode: "; timeout 300 python tools/bench_attn.py $mode 2>&1 | tail -1; done 2>&1 | tee gpurun_out/attn_exp4.txt
MMT_B200_DEV_LIB=ax16 python tools/attn_stamps.py sym 2>&1 | tee gpurun_out/attn_stamps2.txt
