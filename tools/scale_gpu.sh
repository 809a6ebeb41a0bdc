#!/bin/bash
# one multi-GPU bench run the way the driver launches it: tools/scale_gpu.sh <N>
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/scale_${N}gpu.json 2> gpurun_out/scale_${N}gpu.err; echo "bench $N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_${N}gpu.json") if l.startswith("{")][-1])
print(d["n_gpus"], round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms e2e", round(d["e2e"]["value"],1), d["clocks"], d.get("scaling"))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/scale_ref_${N}gpu.json 2> gpurun_out/scale_ref_${N}gpu.err; echo "ref $N rc=$?"; tail -c 300 gpurun_out/scale_ref_${N}gpu.json | head -c 200
