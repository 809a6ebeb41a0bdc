"""CPU: the oracle restatement against the committed golden vectors that oracle/gen_golden.py produced by
running the UNMODIFIED reference (this is what pins the oracle outside the build container)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = ["mixformer_vit", "mixformer_vit_rgbt_shared", "asymmetric_shared_ce"]   # one per backbone family (CPU time)


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_matches_reference_vectors(variant):
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model, cfg = synthetic.make_model(variant, 0)
    inputs = synthetic.make_inputs(variant, cfg, 2, 1)
    out = O.forward(variant, model.state_dict(), cfg, *inputs)
    g = np.load(os.path.join(GOLDEN, f"{variant}_b2.npz"))
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() <= 1e-5
    assert np.abs(out["score_maps"].numpy() - g["score_maps"]).max() <= 2e-4
    if variant == "asymmetric_shared_ce":
        for j in range(3):
            assert np.array_equal(out["ce_keep_v"][j][:2].numpy().astype(np.int32), g[f"ce_keep_v_{j}"])
            assert np.array_equal(out["ce_keep_i"][j][:2].numpy().astype(np.int32), g[f"ce_keep_i_{j}"])
            assert out["ce_keep_v"][j].shape[1] == (227, 159, 112)[j]     # 324 -> 227 -> 159 -> 112


def test_asymmetric_shared_online_oracle_matches_reference_vectors():
    """asymmetric_shared + SPM on the fused map (asymmetric_shared_online.py:337-413): boxes, maps and score logits the
    UNMODIFIED reference module produced (oracle/gen_golden.py main_asym_online)."""
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    variant = "asymmetric_shared_online"
    model, cfg = synthetic.make_model(variant, 0)
    inputs = synthetic.make_inputs(variant, cfg, 2, 1)
    out = O.forward(variant, model.state_dict(), cfg, *inputs)
    g = np.load(os.path.join(GOLDEN, f"{variant}__spm_b2.npz"))
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() <= 1e-5
    assert np.abs(out["score_maps"].numpy() - g["score_maps"]).max() <= 2e-4
    assert np.abs(out["pred_scores"].numpy() - g["pred_scores"]).max() <= 1e-5


@pytest.mark.parametrize("variant", ["mixformer_vit_online", "mixformer_convmae_online"])
def test_online_oracle_matches_reference_vectors(variant):
    """Online trackers: full forward with the SPM score, and set_online + forward_test on seeded crops."""
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model, cfg = synthetic.make_model(variant, 0, sharpen=False)
    sd = model.state_dict()
    g = np.load(os.path.join(GOLDEN, f"{variant}_plain_b2.npz"))
    out = O.forward(variant, sd, cfg, *synthetic.make_inputs(variant, cfg, 2, 1))
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() <= 1e-5
    assert np.abs(out["pred_scores"].numpy() - g["pred_scores"]).max() <= 1e-4
    tt, oo, ss = synthetic.make_online_inputs(cfg, 3, 11)
    st = O.online_set(sd, cfg, tt, oo)
    out2 = O.online_forward_test(sd, cfg, st, ss)
    assert np.abs(out2["pred_boxes"].numpy() - g["online_pred_boxes"]).max() <= 1e-5
    assert np.abs(out2["pred_scores"].numpy() - g["online_pred_scores"]).max() <= 1e-4
    assert np.abs(out2["score_maps"].numpy() - g["online_score_maps"]).max() <= 2e-4


def test_golden_boxes_are_not_degenerate():
    """The sharpened seeded weights must give boxes away from the crop centre, otherwise the 0.5 px bound
    would hold for any implementation ('parity at random init')."""
    for f in sorted(os.listdir(GOLDEN)):
        if "_plain_" in f or "__" in f or f.startswith("frames_"):
            continue      # default-init weight set: flat maps by construction (used for the bf16 tolerances)
        g = np.load(os.path.join(GOLDEN, f))
        cxcy = g["pred_boxes"].reshape(-1, 4)[:, :2] * 288
        assert np.abs(cxcy - 144).max() > 3.0, f
        p = np.exp(g["score_maps"] - g["score_maps"].max(-1, keepdims=True))
        p /= p.sum(-1, keepdims=True)
        assert p.max() > 5.0 / p.shape[-1], f      # peaky, not uniform


def test_msda_core_matches_grid_sample_kat():
    """Restatement of the reference's own MSDA check (deformable_attention/ops/test.py:31-60): the sampling core
    equals F.grid_sample(bilinear, zeros, align_corners=False) at the test's toy shapes, in fp64."""
    import torch.nn.functional as F
    from oracle import mixformer_oracle as O
    torch.manual_seed(3)
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    value = torch.rand(N, S, M, D, dtype=torch.float64) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2, dtype=torch.float64) * 1.4 - 0.2     # includes out-of-map samples
    attn = torch.rand(N, Lq, M, L, P, dtype=torch.float64) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out = O.msda_core(value, shapes, loc, attn)
    # grid_sample formulation
    vals, start = [], 0
    grids = 2 * loc - 1
    for l, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].flatten(2).transpose(1, 2).reshape(N * M, D, H, W)
        start += H * W
        gl = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)
        vals.append(F.grid_sample(v, gl, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = attn.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    ref = (torch.stack(vals, dim=-2).flatten(-2) * aw).sum(-1).view(N, M * D, Lq).transpose(1, 2)
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-12)


def test_forced_keep_mode_reproduces_the_free_run():
    """The oracle's forced-keep test aid (candidate_elimination(..., forced=...)): forcing the kept sets the free run
    chose itself must reproduce it exactly; forcing a different (swapped) token must change the population the next
    stage scores."""
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    variant = "asymmetric_shared_ce"
    model, cfg = synthetic.make_model(variant, 0, sharpen=False)
    sd = model.state_dict()
    inputs = synthetic.make_inputs(variant, cfg, 1, 4)
    free = O.forward(variant, sd, cfg, *inputs)
    forced = [(kv.clone(), ki.clone()) for kv, ki in zip(free["ce_keep_v"], free["ce_keep_i"])]
    again = O.forward(variant, sd, cfg, *inputs, forced_keep=forced)
    assert torch.equal(free["pred_boxes"], again["pred_boxes"]) and torch.equal(free["score_maps"], again["score_maps"])
    for a, b in zip(free["ce_scores"], again["ce_scores"]):
        assert torch.equal(a, b)
    # swap a stage-0 RGB token that the free run prunes at stage 1 for one it removed at stage 0: stage 1 then scores a
    # different population (and the later forced lists stay valid)
    keep0, keep1 = forced[0][0][0].tolist(), set(free["ce_keep_v"][1][0].tolist())
    pos = next(i for i, t in enumerate(keep0) if t not in keep1)
    other = next(float(t) for t in range(324) if float(t) not in set(keep0))
    forced[0][0][0, pos] = other
    diff = O.forward(variant, sd, cfg, *inputs, forced_keep=forced)
    assert other in set(diff["ce_gidx_in_v"][1][0].tolist())
    assert not torch.equal(diff["ce_scores"][1], free["ce_scores"][1])


@pytest.mark.parametrize("yaml_name", ["attention-lasher-cross_deform_fusion_sum_2layer", "attention_lasher_newfusionAdd_2layer"])
def test_oracle_matches_reference_vectors_sum_fusion_classes(yaml_name):
    """Attention_Fusion_Bimodal_LNSpecific_Sum / _2 (fusion_utils.py:282-353): oracle against the reference's outputs."""
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    variant = "asymmetric_shared"
    model, cfg = synthetic.make_model(variant, 0, yaml_name=yaml_name)
    inputs = synthetic.make_inputs(variant, cfg, 2, 1)
    out = O.forward(variant, model.state_dict(), cfg, *inputs)
    g = np.load(os.path.join(GOLDEN, f"{variant}__{yaml_name}_b2.npz"))
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() <= 1e-5
    assert np.abs(out["score_maps"].numpy() - g["score_maps"]).max() <= 2e-4


def test_oracle_matches_reference_vectors_plain_corner_head():
    """HEAD_TYPE = CORNER (Corner_Predictor head.py:23-94): oracle against the reference's outputs."""
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    model, cfg = synthetic.make_model("mixformer_vit", 0, overrides={"MODEL.HEAD_TYPE": "CORNER"})
    inputs = synthetic.make_inputs("mixformer_vit", cfg, 2, 1)
    out = O.forward("mixformer_vit", model.state_dict(), cfg, *inputs)
    g = np.load(os.path.join(GOLDEN, "mixformer_vit__head_corner_b2.npz"))
    assert np.abs(out["pred_boxes"].numpy() - g["pred_boxes"]).max() <= 1e-5
    assert np.abs(out["score_maps"].numpy() - g["score_maps"]).max() <= 2e-4
