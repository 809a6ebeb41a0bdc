#!/bin/bash
# round 2 ncu evidence (one gpurun call, ncu only after the same command exited 0 without it):
#  (1) launch list of a bs=64 step with per-launch duration, DRAM bytes and tensor-pipe activity  -> whole-step tensor-pipe figure
#  (2) launch list of a bs=1 forward (the latency path)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2p_build.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
python tools/bs1_forward.py mixformer_vit_rgbt 3 64 > gpurun_out/r2p_plain64.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2p_step64.csv python tools/bs1_forward.py mixformer_vit_rgbt 3 64 > gpurun_out/r2p_ncu64.log 2>&1
echo "ncu64 rc=$?"; cat gpurun_out/r2p_plain64.log | tail -1
python tools/bs1_forward.py mixformer_vit_rgbt 4 1 > gpurun_out/r2p_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_bs1.csv python tools/bs1_forward.py mixformer_vit_rgbt 4 1 > gpurun_out/r2p_ncu1.log 2>&1
echo "ncu1 rc=$?"; cat gpurun_out/r2p_plain1.log | tail -1
N64=$(grep -o "launches per forward: [0-9]*" gpurun_out/r2p_plain64.log | grep -o "[0-9]*$")
N1=$(grep -o "launches per forward: [0-9]*" gpurun_out/r2p_plain1.log | grep -o "[0-9]*$")
echo "launches per forward: bs64 $N64 bs1 $N1"
python tools/ncu_launch_table.py gpurun_out/r2p_step64.csv > gpurun_out/r2p_step64_table.md 2>&1; tail -25 gpurun_out/r2p_step64_table.md
python tools/ncu_launch_table.py gpurun_out/r2p_bs1.csv > gpurun_out/r2p_bs1_table.md 2>&1; tail -25 gpurun_out/r2p_bs1_table.md
# (3) ncu --set full of the qkv GEMM with the folded-LayerNorm epilogue and of the plain one (source-level stall reasons)
python tools/gemm_case.py qkv ln > gpurun_out/r2p_plain_qkv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2p_qkv_ln python tools/gemm_case.py qkv ln > gpurun_out/r2p_ncu_qkv_ln.log 2>&1
python tools/gemm_case.py qkv plain > gpurun_out/r2p_plain_qkv2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2p_qkv_plain python tools/gemm_case.py qkv plain > gpurun_out/r2p_ncu_qkv_plain.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
