#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2j_build.log 2>&1
for dbg in 0 2048 0 2048; do
  echo "== dev lib MMT_GEMM_DBG=$dbg (2048 = single bf16 staging tile, as before)" >> gpurun_out/r2j_gemm_ln.txt
  MMT_B200_DEV_LIB=1 MMT_GEMM_DBG=$dbg timeout 300 python tools/bench_gemm_ln.py 28928 2>&1 | grep -E "qkv|fc1" >> gpurun_out/r2j_gemm_ln.txt
done
cat gpurun_out/r2j_gemm_ln.txt
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
MMT_LN_FOLD=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2j_bench.json"))
print(round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "bs1 p50", round(d["latency_bs1"]["device_p50_ms"],3), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["share_of_step"],3), d["clocks"])
PY
