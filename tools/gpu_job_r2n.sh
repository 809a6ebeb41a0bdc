#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2n_build.log 2>&1
timeout 300 python tools/bench_gemm_ln.py 28928 2>&1 | grep -v "^  \|Traceback\|File" > gpurun_out/r2n_gemm_ln.txt; cat gpurun_out/r2n_gemm_ln.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n_pytest.log
for cfg in "1 1" "0 1" "1 1"; do
  set -- $cfg
  MMT_LN_FOLD=$1 MMT_PDL=$2 timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-variants --no-frame-path --cpu-budget 0 \
     > gpurun_out/r2n_bench_f$1_p$2.json 2> gpurun_out/r2n_bench_f$1_p$2.err; echo "bench fold=$1 pdl=$2 rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2n_bench_f$1_p$2.json"))
print("fold=$1 pdl=$2", round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "bs1 p50", round(d["latency_bs1"]["device_p50_ms"],3), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["share_of_step"],3), "launches", d["gpu_launches"], d["clocks"])
PY
done
