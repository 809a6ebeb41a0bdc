"""bf16-path error against the reference goldens per variant and weight set, with the LayerNorm fold on and off
(MMT_LN_FOLD): boxes in px, corner score maps abs.  Evidence for DESIGN.md's bf16 error budget."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mmt_b200  # noqa
from mmt_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [(v, None, s) for v in ("mixformer_vit", "mixformer_vit_rgbt", "mixformer_vit_rgbt_shared", "mixformer_vit_rgbt_unibackbone",
                                "asymmetric_shared") for s in (False, True)]
CASES += [("mixformer_vit", "baseline_large", True), ("mixformer_vit_rgbt_shared", "baseline_attention_lasher_newfusion_2layer", True),
          ("asymmetric_shared", "attention_lasher_newfusionAdd_2layer", True)]


def main():
    print("| variant | yaml | weights | fold | boxes px | maps abs | maps / logit range |")
    print("|---|---|---|---|---|---|---|")
    for variant, yaml_name, sharpen in CASES:
        if yaml_name is None:
            g = np.load(os.path.join(GOLDEN, f"{variant}{'' if sharpen else '_plain'}_b2.npz"))
        else:
            g = np.load(os.path.join(GOLDEN, f"{variant}__{yaml_name}_b2.npz"))
        for fold in ("0", "1"):
            os.environ["MMT_LN_FOLD"] = fold
            model, cfg = synthetic.make_model(variant, 0, sharpen=sharpen, yaml_name=yaml_name)
            model = model.cuda().set_precision("bf16")
            inputs = synthetic.make_inputs(variant, cfg, 2, 1, device="cuda")
            res = model.engine().forward(*inputs)
            torch.cuda.synchronize()
            d_box = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * cfg.DATA.SEARCH.SIZE
            d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
            print(f"| {variant} | {yaml_name or 'default'} | {'sharpened' if sharpen else 'plain'} | {fold} | {d_box:.3f} | "
                  f"{d_map:.3e} | {d_map / np.abs(g['score_maps']).max():.2e} |", flush=True)
            del model
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
