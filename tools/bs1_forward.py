"""A few forwards of one variant at a given batch (eager launches on the engine's streams), for ncu launch lists:
    python tools/bs1_forward.py <variant> <n forwards> <batch> [yaml name]
(batch 1 = the latency path; batch 64 = one bench step per forward; yaml name e.g. baseline_large for the -L models)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import synthetic

variant = sys.argv[1] if len(sys.argv) > 1 else "mixformer_vit_rgbt"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
yaml_name = sys.argv[4] if len(sys.argv) > 4 else None
model, cfg = synthetic.make_model(variant, 0, **({"yaml_name": yaml_name} if yaml_name else {}))
model = model.cuda()
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
inputs = synthetic.make_inputs(variant, cfg, batch, 99, device="cuda")
for _ in range(n):
    out, boxes = model(*inputs)
torch.cuda.synchronize()
print(boxes.view(-1)[:4].tolist(), "launches per forward:", __import__("mmt_b200").ops.LAUNCHES // n)
