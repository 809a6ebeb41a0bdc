#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2q_build.log 2>&1
timeout 300 python tools/bench_gemm_ln.py 28928 2>&1 | grep -E "^(qkv|fc1|proj|fc2)" | tee gpurun_out/r2q_gemm_ln.txt
timeout 120 python tools/bench_attn.py sym; timeout 120 python tools/bench_attn.py cross
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/r2q_bench$i.json 2> gpurun_out/r2q_bench$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2q_bench$i.json"))
print(round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "bs1 p50", round(d["latency_bs1"]["device_p50_ms"],3), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["share_of_step"],3), d["clocks"])
PY
done
