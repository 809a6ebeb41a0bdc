#!/bin/bash
# full GPU validation of the shipped build: every GPU test, smoke(), the default bench line and the reference arm (one gpurun call)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2y_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2y_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_ref.json 2> gpurun_out/r2y_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2y_bench.json"))
print(round(d["value"],1), "frames/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "bs1", d["latency_bs1"]["device_p50_ms"], d["latency_bs1"]["e2e_host_p50_ms"], "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), round(d["roofline"]["share_of_step"],3), "nominal frac", round(d["step_tensor_frac_of_nominal_2250"],4), d["clocks"], "launches", d["gpu_launches"])
print("cached", d["cached_template"]["value"], "frame_path", d["frame_path"]["value"], "e2e fp32 crops", d["e2e_fp32_crops"]["value"])
for k,v in d["configs"].items(): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","ms_per_step","frac_of_nominal_2250","error")})
for k,v in d["gpu_eager_baseline"]["modes"].items(): print("eager",k,{a:round(b,2) for a,b in v.items() if isinstance(b,float)})
print("cpu", d["cpu_baseline"])
r=json.load(open("gpurun_out/r2y_ref.json")); print("ref arm", r["value"], r["steps"], r["warmup"], r["cpu_baseline"]["sample"][:80], r["config"]==d["config"])
PY
