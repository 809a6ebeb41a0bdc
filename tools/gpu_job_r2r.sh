#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2r_build.log 2>&1
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/r2r_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -3 gpurun_out/r2r_pytest_gemm.log
for cl in 1 0 1 0; do
  echo "== MMT_CL4=$cl"; MMT_CL4=$cl timeout 200 python tools/bench_gemm_ln.py 28928 2>&1 | grep -E "^(qkv|fc1|proj|fc2)"
done 2>&1 | tee gpurun_out/r2r_gemm_ln.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
for cl in 1 0 1; do
MMT_CL4=$cl timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 > gpurun_out/r2r_bench_cl$cl.json 2> gpurun_out/r2r_bench_cl$cl.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2r_bench_cl$cl.json"))
print("cl4=$cl", round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "bs1 p50", round(d["latency_bs1"]["device_p50_ms"],3), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["share_of_step"],3), d["clocks"])
PY
done
