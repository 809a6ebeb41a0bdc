#!/bin/bash
# Developer tool (run under gpurun, one GPU): final launch list of the round + ncu --set full of the persistent attention
# kernel and the 8-lane MSDA kernel.  Every ncu command runs only after the same command exited 0 without ncu.
set -u
B="python bench.py --steps 2 --warmup 1 --cpu-budget 0 --no-latency"
$B > gpurun_out/plain_r1d.log 2> gpurun_out/plain_r1d.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1800 -c 900 \
    --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/ncu_r1d.log 2>&1
python tools/bench_attn.py sym short > gpurun_out/p_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc_persist -s 2 -c 1 -f -o gpurun_out/prof_r1d_attn_persist \
    python tools/bench_attn.py sym short > gpurun_out/n_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msda_bimodal_bf16_p4 -s 2 -c 1 -f -o gpurun_out/prof_r1d_msda \
    $B > gpurun_out/n_msda.log 2>&1
ls -la gpurun_out/*r1d*
