#!/bin/bash
# attention timing experiments (developer builds csrc/libmmt_b200_ax*.so; results wrong by construction, timing only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3
for v in "" ax1 ax2 ax3; do
  for mode in sym cross; do
    echo -n "variant '$v' $mode: "; MMT_B200_DEV_LIB=$v timeout 300 python tools/bench_attn.py $mode 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/attn_exp2.txt
timeout 900 python -m pytest tests/test_forward_gpu.py -x -q -k "bf16_mode or template_cache or batch_against" 2>&1 | tail -3
