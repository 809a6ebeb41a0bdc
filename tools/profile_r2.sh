#!/bin/bash
# (final build of round 2: the tables of profiles/r2_launches.md)
# round 2 ncu evidence (one gpurun call, ncu only after the same command exited 0 without it):
#  (1) launch list of a bs=64 step with per-launch duration, DRAM bytes and tensor-pipe activity  -> whole-step tensor-pipe figure
#  (2) launch list of a bs=1 forward (the latency path)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2v_build.log 2>&1
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_native_ops_gpu.py -x -q -k "fusion or in_place or bf16_mode or fp32_mode" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python -c "import json;d=json.load(open('gpurun_out/r2v_bench.json'));print(round(d['value'],1), round(d['ms_per_step'],3), 'bs1', d['latency_bs1']['device_p50_ms'], d['latency_bs1']['e2e_host_p50_ms'], 'e2e', round(d['e2e']['value'],1))"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
python tools/bs1_forward.py mixformer_vit_rgbt 3 64 > gpurun_out/r2v_plain64.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2v_step64.csv python tools/bs1_forward.py mixformer_vit_rgbt 3 64 > gpurun_out/r2v_ncu64.log 2>&1
echo "ncu64 rc=$?"; cat gpurun_out/r2v_plain64.log | tail -1
python tools/bs1_forward.py mixformer_vit_rgbt 4 1 > gpurun_out/r2v_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_bs1.csv python tools/bs1_forward.py mixformer_vit_rgbt 4 1 > gpurun_out/r2v_ncu1.log 2>&1
echo "ncu1 rc=$?"; cat gpurun_out/r2v_plain1.log | tail -1
N64=$(grep -o "launches per forward: [0-9]*" gpurun_out/r2v_plain64.log | grep -o "[0-9]*$")
N1=$(grep -o "launches per forward: [0-9]*" gpurun_out/r2v_plain1.log | grep -o "[0-9]*$")
echo "launches per forward: bs64 $N64 bs1 $N1"
python tools/ncu_launch_table.py gpurun_out/r2v_step64.csv > gpurun_out/r2v_step64_table.md 2>&1; tail -25 gpurun_out/r2v_step64_table.md
python tools/ncu_launch_table.py gpurun_out/r2v_bs1.csv > gpurun_out/r2v_bs1_table.md 2>&1; tail -25 gpurun_out/r2v_bs1_table.md
