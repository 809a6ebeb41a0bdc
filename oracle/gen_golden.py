"""TEST INFRASTRUCTURE - pins the oracle against the UNMODIFIED reference and writes golden vectors.

Runs only in the build container (needs /root/reference).  For every variant on the accelerated path:
  1. build the mmt_b200 model with seeded, sharpened weights (multi-modal-tracking_b200/synthetic.py);
  2. build the reference model from its own shipped YAML (oracle/ref_shims.py) and load OUR state_dict into it
     with strict=True  -> pins the checkpoint key/shape layout of the drop-in builders;
  3. run the reference forward and oracle/mixformer_oracle.forward on the same seeded N(0,1) inputs (CPU fp32)
     and require agreement to float round-off  -> pins the oracle;
  4. save the REFERENCE's outputs (boxes, corner score maps via a forward hook on the head, CE kept indices via a
     hook on the CE blocks) to tests/golden/<variant>_b<B>.npz.
Usage:  python oracle/gen_golden.py [variant ...]
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mmt_b200  # noqa: E402,F401
from mmt_b200 import synthetic  # noqa: E402
from oracle import mixformer_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
BATCH = 2
WEIGHT_SEED, INPUT_SEED = 0, 1


def run_reference(variant, sd, inputs):
    model, rcfg = ref_shims.build_reference_model(variant, synthetic.DEFAULT_YAML[variant])
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    cap = {}
    head = model.box_head
    orig = head.get_score_map

    def hooked(x):
        tl, br = orig(x)
        cap["maps"] = torch.stack([tl.flatten(1), br.flatten(1)], dim=1)
        cap["feat"] = x
        return tl, br

    head.get_score_map = hooked
    keeps_v, keeps_i = [], []
    if variant == "asymmetric_shared_ce":
        for i, blk in enumerate(model.backbone.blocks):
            if i in rcfg.MODEL.BACKBONE.CE_LOC:
                def hook(m, a, out):
                    keeps_v.append(out[2].clone())
                    keeps_i.append(out[3].clone())
                blk.register_forward_hook(hook)
    with torch.no_grad():
        out, coords = model(*inputs)
    res = dict(pred_boxes=out["pred_boxes"], score_maps=cap["maps"], feat=cap["feat"])
    if keeps_v:
        res["ce_keep_v"], res["ce_keep_i"] = keeps_v, keeps_i
    return res


def _cpu_prroi_module():
    """PrRoIPool2D is GPU-only in the reference (prroi_pool/functional.py:62-63); on CPU the reference's ScoreDecoder
    is run with the oracle restatement of the pooling plugged in (itself pinned by the reference's known-answer test)."""
    from oracle import native_ops_oracle as NO

    class P(torch.nn.Module):
        def forward(self, features, rois):
            return torch.from_numpy(NO.prroi_pool_forward(features.numpy(), rois.numpy(), 4, 4, 1.0))
    return P()


def main_online(sharpen, yaml_name=None, batch=BATCH, variant="mixformer_vit_online"):
    """mixformer_vit_online / mixformer_convmae_online: full forward with the SPM score head, and set_online +
    forward_test (cached templates).  yaml_name = "baseline_large": the -L models (384 search / 192 template), batch 1."""
    yaml_name = yaml_name or synthetic.DEFAULT_YAML[variant]
    model, cfg = synthetic.make_model(variant, WEIGHT_SEED, sharpen=sharpen, yaml_name=yaml_name)
    sd = model.state_dict()
    ref, rcfg = ref_shims.build_reference_model(variant, yaml_name)
    missing, unexpected = ref.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    ref.score_branch.search_prroipool = _cpu_prroi_module()
    cap = {}
    orig = ref.box_head.get_score_map

    def hooked(x):
        tl, br = orig(x)
        cap["maps"] = torch.stack([tl.flatten(1), br.flatten(1)], dim=1)
        return tl, br
    ref.box_head.get_score_map = hooked
    save = {}
    t, ot, s = synthetic.make_inputs(variant, cfg, batch, INPUT_SEED)
    with torch.no_grad():
        out, _ = ref(t, ot, s, run_score_head=True)
    ora = O.forward(variant, sd, cfg, t, ot, s)
    d = [(out["pred_boxes"] - ora["pred_boxes"]).abs().max().item(), (cap["maps"] - ora["score_maps"]).abs().max().item(),
         (out["pred_scores"] - ora["pred_scores"]).abs().max().item()]
    print(f"{variant} ({'sharpened' if sharpen else 'plain'}) full: oracle vs reference boxes {d[0]:.3e} maps {d[1]:.3e} scores {d[2]:.3e}")
    assert d[0] <= 1e-5 and d[1] <= 2e-4 and d[2] <= 2e-4
    save.update(pred_boxes=out["pred_boxes"].numpy(), score_maps=cap["maps"].numpy(), pred_scores=out["pred_scores"].numpy())
    # cached-template path: one template, 3 online templates, one search crop
    tt, oo, ss = synthetic.make_online_inputs(cfg, 3, INPUT_SEED + 10)
    with torch.no_grad():
        ref.set_online(tt, oo)
        out2, _ = ref.forward_test(ss, run_score_head=True)
    st = O.online_set(sd, cfg, tt, oo)
    ora2 = O.online_forward_test(sd, cfg, st, ss)
    d = [(out2["pred_boxes"] - ora2["pred_boxes"]).abs().max().item(), (cap["maps"] - ora2["score_maps"]).abs().max().item(),
         (out2["pred_scores"] - ora2["pred_scores"]).abs().max().item()]
    print(f"   cached-template path: oracle vs reference boxes {d[0]:.3e} maps {d[1]:.3e} scores {d[2]:.3e}")
    assert d[0] <= 1e-5 and d[1] <= 2e-4 and d[2] <= 2e-4
    save.update(online_pred_boxes=out2["pred_boxes"].numpy(), online_score_maps=cap["maps"].numpy(),
                online_pred_scores=out2["pred_scores"].numpy())
    print("   scores (logits):", out["pred_scores"].tolist(), out2["pred_scores"].tolist())
    tag = ("" if sharpen else "_plain") + ("" if yaml_name == synthetic.DEFAULT_YAML[variant] else "_" + yaml_name)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{variant}{tag}_b{batch}.npz"), **save)


EXTRA = [("mixformer_vit_rgbt_shared", "baseline_attention_lasher_newfusion_2layer"),   # Attention_Fusion_Bimodal
         ("asymmetric_shared", "attention_lasher_cat_3layer"),                           # RGBT_Fusion_Cat
         ("mixformer_vit", "baseline_large"),                                            # MixViT-L RGB-only, 384 / 192
         # Attention_Fusion_Bimodal_LNSpecific_Sum / _2 (fusion_utils.py:282-353), the fusion classes of three shipped YAMLs
         ("asymmetric_shared", "attention-lasher-cross_deform_fusion_sum_2layer"),
         ("asymmetric_shared", "attention_lasher_newfusionAdd_2layer"),
         ("asymmetric_shared_ce", "attention_lasher_newfusionAdd_2layer")]


def main_asym_online():
    """asymmetric_shared_online (asymmetric_shared + SPM on the fused map, asymmetric_shared_online.py:337-413): the
    reference module with run_score_head=True against the oracle; sharpened and plain weight sets, batch 2."""
    variant = "asymmetric_shared_online"
    for sharpen in (True, False):
        model, cfg = synthetic.make_model(variant, WEIGHT_SEED, sharpen=sharpen)
        sd = model.state_dict()
        inputs = synthetic.make_inputs(variant, cfg, BATCH, INPUT_SEED)
        ref_model, _ = ref_shims.build_reference_model(variant, synthetic.DEFAULT_YAML[variant])
        missing, unexpected = ref_model.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        ref_model.score_branch.search_prroipool = _cpu_prroi_module()
        cap = {}
        orig = ref_model.box_head.get_score_map

        def hooked(x, orig=orig, cap=cap):
            tl, br = orig(x)
            cap["maps"] = torch.stack([tl.flatten(1), br.flatten(1)], dim=1)
            return tl, br
        ref_model.box_head.get_score_map = hooked
        with torch.no_grad():
            out, _ = ref_model(*inputs, run_score_head=True)
        ora = O.forward(variant, sd, cfg, *inputs)
        d_box = (out["pred_boxes"] - ora["pred_boxes"]).abs().max().item()
        d_map = (cap["maps"] - ora["score_maps"]).abs().max().item()
        d_sc = (out["pred_scores"] - ora["pred_scores"]).abs().max().item()
        print(f"{variant} ({'sharpened' if sharpen else 'plain'}): oracle vs reference boxes {d_box:.3e} maps {d_map:.3e} "
              f"scores {d_sc:.3e}   scores {out['pred_scores'].tolist()}")
        assert d_box <= 1e-5 and d_map <= 2e-4 and d_sc <= 1e-4
        tag = "" if sharpen else "_plain"
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{variant}__spm{tag}_b{BATCH}.npz"),
                            pred_boxes=out["pred_boxes"].numpy(), score_maps=cap["maps"].numpy(),
                            pred_scores=out["pred_scores"].numpy())


ONLY = []      # restrict main_extra to these (variant, yaml) pairs (command line: extra:<variant>:<yaml>)


def main_extra():
    """The two remaining fusion classes of the shipped YAMLs (sharpened weights, batch 2)."""
    for variant, yaml_name in EXTRA:
        if ONLY and (variant, yaml_name) not in ONLY:
            continue
        model, cfg = synthetic.make_model(variant, WEIGHT_SEED, yaml_name=yaml_name)
        sd = model.state_dict()
        inputs = synthetic.make_inputs(variant, cfg, BATCH, INPUT_SEED)
        ref_model, rcfg = ref_shims.build_reference_model(variant, yaml_name)
        missing, unexpected = ref_model.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        cap = {}
        orig = ref_model.box_head.get_score_map

        def hooked(x, orig=orig, cap=cap):
            tl, br = orig(x)
            cap["maps"] = torch.stack([tl.flatten(1), br.flatten(1)], dim=1)
            return tl, br
        ref_model.box_head.get_score_map = hooked
        with torch.no_grad():
            out, _ = ref_model(*inputs)
        ora = O.forward(variant, sd, cfg, *inputs)
        d_box = (out["pred_boxes"] - ora["pred_boxes"]).abs().max().item()
        d_map = (cap["maps"] - ora["score_maps"]).abs().max().item()
        print(f"{variant}/{yaml_name} ({cfg.MODEL.get('FUSION_CLASS')}): oracle vs reference boxes {d_box:.3e} maps {d_map:.3e}")
        assert d_box <= 1e-5 and d_map <= 2e-4
        fp32_maps = cap["maps"]
        # yardstick for the bf16 tests on this stress set: the UNMODIFIED reference in its own reduced-precision mode
        # (torch.autocast(bfloat16) on the CPU) - how far the reference's boxes move when only the precision changes
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            out_ac, _ = ref_model(*inputs)
        ac_maps = cap["maps"].float()
        print(f"  reference under autocast(bf16): boxes move {(out_ac['pred_boxes'].float() - out['pred_boxes']).abs().max().item() * cfg.DATA.SEARCH.SIZE:.3f} px, "
              f"maps {(ac_maps - fp32_maps).abs().max().item():.3e}")
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{variant}__{yaml_name}_b{BATCH}.npz"),
                            pred_boxes=out["pred_boxes"].numpy(), score_maps=fp32_maps.numpy(),
                            pred_boxes_autocast_bf16=out_ac["pred_boxes"].float().numpy(),
                            score_maps_autocast_bf16=ac_maps.numpy())


def main_corner_head():
    """The plain Corner_Predictor (HEAD_TYPE = CORNER, lib/models/mixformer_cvt/head.py:23-94) behind the MixViT-B backbone:
    the reference supports it through build_box_head (head.py:235-258) but ships MixViT YAMLs with CORNER_UP only, so the
    reference model is built from experiments/mixformer_vit/baseline.yaml with MODEL.HEAD_TYPE overridden."""
    variant, ov = "mixformer_vit", {"MODEL.HEAD_TYPE": "CORNER"}
    model, cfg = synthetic.make_model(variant, WEIGHT_SEED, overrides=ov)
    sd = model.state_dict()
    inputs = synthetic.make_inputs(variant, cfg, BATCH, INPUT_SEED)
    ref_model, _ = ref_shims.build_reference_model(variant, synthetic.DEFAULT_YAML[variant], overrides=ov)
    missing, unexpected = ref_model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    cap = {}
    orig = ref_model.box_head.get_score_map

    def hooked(x):
        tl, br = orig(x)
        cap["maps"] = torch.stack([tl.flatten(1), br.flatten(1)], dim=1)
        return tl, br
    ref_model.box_head.get_score_map = hooked
    with torch.no_grad():
        out, _ = ref_model(*inputs)
    ora = O.forward(variant, sd, cfg, *inputs)
    d_box = (out["pred_boxes"] - ora["pred_boxes"]).abs().max().item()
    d_map = (cap["maps"] - ora["score_maps"]).abs().max().item()
    print(f"{variant} HEAD_TYPE=CORNER: oracle vs reference boxes {d_box:.3e} maps {d_map:.3e}; boxes {out['pred_boxes'].view(-1, 4).tolist()}")
    assert d_box <= 1e-5 and d_map <= 2e-4
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{variant}__head_corner_b{BATCH}.npz"),
                        pred_boxes=out["pred_boxes"].numpy(), score_maps=cap["maps"].numpy())


def main(variants):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if "corner_head" in variants:
        torch.set_num_threads(8)
        main_corner_head()
        variants = [v for v in variants if v != "corner_head"]
    for v in [v for v in variants if v.startswith("extra:")]:
        ONLY.append(tuple(v.split(":")[1:3]))
        variants = [x for x in variants if x != v] + (["extra"] if "extra" not in variants else [])
    if "asymmetric_shared_online" in variants:
        torch.set_num_threads(8)
        main_asym_online()
        variants = [v for v in variants if v != "asymmetric_shared_online"]
    if "extra" in variants:
        torch.set_num_threads(8)
        main_extra()
        variants = [v for v in variants if v != "extra"]
    torch.set_num_threads(8)
    torch.set_num_threads(8)
    for ov in ("mixformer_vit_online", "mixformer_convmae_online"):
        if ov in variants:
            for sharpen in (True, False):
                main_online(sharpen, variant=ov)
            main_online(False, "baseline_large", 1, variant=ov)
            variants = [v for v in variants if v != ov]
    for variant, sharpen in [(v, s) for v in variants for s in (True, False)]:
        # two seeded weight sets: "sharpened" (peaky corner maps, see synthetic.py) and "plain" = the builders'
        # default random init, the weight set BASELINE.json's north_star names for the bf16 tolerances
        model, cfg = synthetic.make_model(variant, WEIGHT_SEED, sharpen=sharpen)
        sd = model.state_dict()
        inputs = synthetic.make_inputs(variant, cfg, BATCH, INPUT_SEED)
        ref = run_reference(variant, sd, inputs)
        ora = O.forward(variant, sd, cfg, *inputs)
        d_box = (ref["pred_boxes"] - ora["pred_boxes"]).abs().max().item()
        d_map = (ref["score_maps"] - ora["score_maps"]).abs().max().item()
        d_feat = (ref["feat"] - ora["feat"]).abs().max().item()
        print(f"{variant} ({'sharpened' if sharpen else 'plain'}): oracle vs reference  boxes {d_box:.3e}  score maps {d_map:.3e}  head input {d_feat:.3e}")
        assert d_box <= 1e-5 and d_map <= 2e-4 and d_feat <= 2e-4, "oracle does not restate the reference"
        save = dict(pred_boxes=ref["pred_boxes"].numpy(), score_maps=ref["score_maps"].numpy(),
                    feat_mean_abs=np.float32(ref["feat"].abs().mean().item()))
        if "ce_keep_v" in ref:
            for k in ("ce_keep_v", "ce_keep_i"):
                for j, (a, b) in enumerate(zip(ref[k], ora[k])):
                    assert torch.equal(a, b[: a.shape[0]]), f"{k}[{j}] differs between oracle and reference"
                    save[f"{k}_{j}"] = a.numpy().astype(np.int32)
            for j, s in enumerate(ora["ce_scores"]):
                save[f"ce_scores_{j}"] = s.numpy()
        px = ref["pred_boxes"].view(-1, 4) * cfg.DATA.SEARCH.SIZE
        print("   reference boxes (px, cxcywh):", np.round(px.numpy(), 2).tolist())
        tag = "" if sharpen else "_plain"
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{variant}{tag}_b{BATCH}.npz"), **save)


if __name__ == "__main__":
    main(sys.argv[1:] or list(synthetic.DEFAULT_YAML))
