#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
for spec in "mixformer_convmae_online 2 32 baseline_large:cmae" "asymmetric_shared_ce 2 128:ce"; do
  args=${spec%%:*}; tag=${spec##*:}
  python tools/bs1_forward.py $args > gpurun_out/r2c_${tag}_plain.log 2>&1 && \
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2c_${tag}.csv python tools/bs1_forward.py $args > gpurun_out/r2c_${tag}_ncu.log 2>&1
  echo "$tag rc=$?"; tail -1 gpurun_out/r2c_${tag}_plain.log
  python tools/ncu_launch_table.py gpurun_out/r2c_${tag}.csv > gpurun_out/r2c_${tag}_table.md 2>&1; grep -v "at::" gpurun_out/r2c_${tag}_table.md | head -24; tail -1 gpurun_out/r2c_${tag}_table.md
done
