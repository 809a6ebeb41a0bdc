#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2s_build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2s_bench_2gpu.json 2> gpurun_out/r2s_bench_2gpu.err; echo "2gpu rc=$?"
tail -c 600 gpurun_out/r2s_bench_2gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2s_bench_2gpu.json"))
print(d["n_gpus"], round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d["clocks"], d["config"]["parallelism"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2s_ref_2gpu.json 2> gpurun_out/r2s_ref_2gpu.err; echo "ref 2gpu rc=$?"; cut -c1-300 gpurun_out/r2s_ref_2gpu.json
