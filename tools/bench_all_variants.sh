#!/bin/bash
# Developer tool: one bench line per covered configuration (BASELINE.json configs 1-5), results into gpurun_out/variants.jsonl
out=gpurun_out/variants.jsonl; : > $out
run() { python bench.py --steps 20 --warmup 3 --cpu-budget 0 --no-latency --no-frame-path "$@" 2>/dev/null | tail -1 >> $out; }
run --variant mixformer_vit --batch 64
run --variant mixformer_vit_rgbt --batch 64
run --variant mixformer_vit_rgbt_shared --batch 64
run --variant mixformer_vit_rgbt_unibackbone --batch 64
run --variant asymmetric_shared --batch 64
run --variant asymmetric_shared_ce --batch 64
run --variant asymmetric_shared_ce --batch 128
run --variant mixformer_vit_online --batch 64
run --variant mixformer_vit_online --yaml baseline_large --batch 32
run --variant mixformer_convmae_online --batch 64
run --variant mixformer_convmae_online --yaml baseline_large --batch 32
python - <<'PY'
import json
for l in open("gpurun_out/variants.jsonl"):
    d = json.loads(l)
    c = d.get("cached_template") or {}
    print(d["config"]["workload"][:70], "| bs", d["config"]["batch_per_gpu"], "|", round(d["value"], 1), "fps |", round(d["ms_per_step"], 2), "ms | e2e",
          round(d["e2e"]["value"], 1), "| gemm", round(d["roofline"]["achieved"], 1), "TF | step frac", round(d["step_tensor_frac_of_measured_peak"], 3),
          "| cached", round(c.get("value", 0), 1))
PY
