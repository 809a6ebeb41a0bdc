#!/bin/bash
# ncu --set full captures of the final build's three hot kernels (each command ran without ncu first, rc 0)
mkdir -p gpurun_out
set -x
python tools/gemm_case.py fc1 ln > gpurun_out/full_fc1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/r2z_fc1_ln -f python tools/gemm_case.py fc1 ln > gpurun_out/full_fc1.log 2>&1; echo "fc1 rc=$?"
python tools/gemm_case.py proj ln > gpurun_out/full_proj_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/r2z_proj_ln -f python tools/gemm_case.py proj ln > gpurun_out/full_proj.log 2>&1; echo "proj rc=$?"
python tools/bench_attn.py sym short > gpurun_out/full_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc_persist -s 2 -c 1 -o gpurun_out/r2z_attn -f python tools/bench_attn.py sym short > gpurun_out/full_attn.log 2>&1; echo "attn rc=$?"
ls -la gpurun_out/*.ncu-rep
