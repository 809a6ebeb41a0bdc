#!/bin/bash
# Developer tool (run under gpurun, one GPU): launch list of the bench command + ncu --set full of the CTA-pair GEMM with
# the fp32 TMA epilogue (fc2 shape) and of the frame-crop kernel.  Every ncu command runs only after the same command
# exited 0 without ncu.  Outputs land in gpurun_out/.
set -u
B="python bench.py --steps 2 --warmup 1 --cpu-budget 0 --no-latency"
$B > gpurun_out/plain_r1c.log 2> gpurun_out/plain_r1c.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1800 -c 900 \
    --csv --log-file gpurun_out/launches_r1c.csv $B > gpurun_out/ncu_r1c.log 2>&1
python tools/bench_gemm.py fc2 > gpurun_out/p_fc2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 3 -c 1 -f -o gpurun_out/prof_r1c_gemm_pair_fc2 \
    python tools/bench_gemm.py fc2 > gpurun_out/n_fc2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:frame_crop_kernel -s 4 -c 1 -f -o gpurun_out/prof_r1c_frame_crop \
    $B > gpurun_out/n_crop.log 2>&1
ls -la gpurun_out/*r1c*
