#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2m_build.log 2>&1
for i in 1 2; do
  echo "== dev lib (poly on)"; MMT_B200_DEV_LIB=1 timeout 120 python tools/bench_attn.py sym; MMT_B200_DEV_LIB=1 timeout 120 python tools/bench_attn.py cross
  echo "== shipped (poly off)"; timeout 120 python tools/bench_attn.py sym; timeout 120 python tools/bench_attn.py cross
done 2>&1 | tee gpurun_out/r2m_attn.txt
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_forward_gpu.py -x -q -k "attention or bf16 or full_size" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2m_bench.json"))
print(round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "bs1 p50", round(d["latency_bs1"]["device_p50_ms"],3), "roof", round(d["roofline"]["achieved"],1), d["clocks"])
PY
