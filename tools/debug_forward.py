"""Developer tool: run each variant twice per precision on the GPU; report determinism and error vs golden."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmt_b200  # noqa
from mmt_b200 import synthetic
GOLDEN = os.path.join(ROOT, "tests", "golden")
variants = sys.argv[1:] or ["mixformer_vit", "mixformer_vit_rgbt", "mixformer_vit_rgbt_shared",
                            "mixformer_vit_rgbt_unibackbone", "asymmetric_shared", "asymmetric_shared_ce"]
for variant in variants:
    g = np.load(os.path.join(GOLDEN, f"{variant}_b2.npz"))
    for prec in ("fp32", "bf16"):
        model, cfg = synthetic.make_model(variant, 0)
        model = model.cuda().set_precision(prec)
        inputs = synthetic.make_inputs(variant, cfg, 2, 1, device="cuda")
        eng = model.engine()
        r1 = eng.forward(*inputs); torch.cuda.synchronize()
        b1, m1, f1 = r1["pred_boxes"].clone(), r1["score_maps"].clone(), r1["feat_rows"].float().clone()
        r2 = eng.forward(*inputs); torch.cuda.synchronize()
        b2, m2, f2 = r2["pred_boxes"], r2["score_maps"], r2["feat_rows"].float()
        print(f"{variant:34s} {prec}: rerun d_box {(b1-b2).abs().max().item():.2e} d_map {(m1-m2).abs().max().item():.2e} "
              f"d_feat {(f1-f2).abs().max().item():.2e} | vs golden: box_px {np.abs(b1.cpu().numpy()-g['pred_boxes']).max()*cfg.DATA.SEARCH.SIZE:.4f} "
              f"map {np.abs(m1.cpu().numpy()-g['score_maps']).max():.3e} (|map| max {np.abs(g['score_maps']).max():.2f})"
              + (f" feat {np.abs(f1.cpu().numpy().reshape(2,-1,f1.shape[-1]) - g['feat'].transpose(0,2,3,1).reshape(2,-1,f1.shape[-1])).max():.3e}" if 'feat' in g else ""), flush=True)
