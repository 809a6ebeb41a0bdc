"""GPU parity of the whole per-frame forward against (a) the committed golden vectors produced by the
UNMODIFIED reference (tests/golden, oracle/gen_golden.py) and (b) the CPU oracle run live on the same inputs.

Thresholds are the north-star ones: boxes <= 0.5 px and score maps <= 1e-2 abs in bf16 mode;
<= 1e-4 (normalised boxes / score maps) in fp32 mode; CE kept-token indices bit-exact in fp32 mode
(swaps allowed only between entries whose oracle scores are closer than 1e-7 relative).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = ["mixformer_vit", "mixformer_vit_rgbt", "mixformer_vit_rgbt_shared", "mixformer_vit_rgbt_unibackbone",
            "asymmetric_shared", "asymmetric_shared_ce"]


def _run(variant, precision, batch=2, sharpen=True):
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(variant, 0, sharpen=sharpen)
    model = model.cuda().set_precision(precision)
    inputs = synthetic.make_inputs(variant, cfg, batch, 1, device="cuda")
    out, coords = model(*inputs)
    torch.cuda.synchronize()
    res = model._engine.forward(*inputs)      # same call, with the auxiliary outputs (score maps, CE indices)
    torch.cuda.synchronize()
    assert torch.equal(res["pred_boxes"], coords), "forward is not run-to-run deterministic"
    assert out["pred_boxes"].shape == (batch, 1, 4)
    return res, cfg


def _golden(variant, sharpen):
    return np.load(os.path.join(GOLDEN, f"{variant}{'' if sharpen else '_plain'}_b2.npz"))


def _ce_stage_matches(keeps_mine, g, j, m, B):
    """Kept global indices of CE stage j, modality m: identical to the reference's, or differing only where the
    reference's own scores of the swapped tokens are closer than 1e-7 relative (CE bit-exactness)."""
    key = ("ce_keep_v", "ce_keep_i")[m]
    ref = g[f"{key}_{j}"]
    mine = keeps_mine[j][m * B:(m + 1) * B]
    if np.array_equal(mine, ref):
        return True
    sc = g[f"ce_scores_{j}"]                       # [B, 2*Ls_j], indexed by the token's LOCAL position at stage j
    Ls = sc.shape[1] // 2
    for b in range(B):
        if j == 0:
            local = {int(t): int(t) for t in range(Ls)}
        else:
            local = {int(t): i for i, t in enumerate(g[f"{key}_{j - 1}"][b])}
        for a, r in zip(mine[b], ref[b]):
            if a != r:
                if int(a) not in local:
                    return False
                sa, sr = sc[b, m * Ls + local[int(a)]], sc[b, m * Ls + local[int(r)]]
                if abs(sa - sr) > 1e-7 * max(abs(sa), abs(sr)):
                    return False
    return True


# Relative distance from the keep boundary inside which a bf16 run may legitimately pick the other token: the CE score of
# a search token is the mean over 2*Lt template rows and 12 heads of softmax probabilities computed from bf16 q/k that
# themselves come out of bf16 GEMMs; the largest gap measured on B200 (bs = 2 goldens, bs = 128 live oracle) is 4.3e-4 of the
# boundary score, the bound leaves 10x margin.  fp32 mode: only float summation order differs.
CE_TIE_REL = {"bf16": 5e-3, "fp32": 1e-6}


def _forced_keep_from(res, rows, B):
    """Per CE stage the (keep_v, keep_i) global-index tensors of the sequences `rows` out of a modality-major
    [2B, keep] engine result."""
    out = []
    for k in res["ce_keep"]:
        k = k.cpu()
        out.append((k[:B][rows].to(torch.float32), k[B:][rows].to(torch.float32)))
    return out


def _assert_ce_keep_sets_are_score_consistent(ora, forced, rel):
    """`ora` = oracle run in forced-keep mode (same token population as the GPU run at every stage).  Every token the GPU
    kept but the oracle's own top-k would not (and vice versa) must lie within `rel` (relative) of the oracle's
    boundary score: a tie-break, not an arithmetic error.  Returns (number of differing tokens, largest relative gap)."""
    n_diff, worst = 0, 0.0
    for j, (kv, ki) in enumerate(forced):
        sc = ora["ce_scores"][j]
        Ls = sc.shape[1] // 2
        for m, (forced_m, gin) in enumerate(((kv, ora["ce_gidx_in_v"][j]), (ki, ora["ce_gidx_in_i"][j]))):
            keep = forced_m.shape[1]
            for b in range(forced_m.shape[0]):
                s_b = sc[b, m * Ls:(m + 1) * Ls]
                pos = {int(g): i for i, g in enumerate(gin[b].tolist())}
                mine = {pos[int(g)] for g in forced_m[b].tolist()}
                order = torch.sort(s_b, descending=True).indices.tolist()
                theirs = set(order[:keep])
                boundary = 0.5 * (s_b[order[keep - 1]] + s_b[order[keep]]).item()
                for i in mine ^ theirs:
                    gap = abs(s_b[i].item() - boundary) / abs(boundary)
                    n_diff += 1
                    worst = max(worst, gap)
                    assert gap <= rel, (f"CE stage {j} modality {m} sequence {b}: token at local position {i} differs from "
                                        f"the oracle's top-{keep} with a score {gap:.3e} (relative) off the boundary")
    return n_diff, worst


def _ce_forced_oracle(variant, model_cpu_sd, cfg, inputs_cpu, res, rows, B):
    from oracle import mixformer_oracle as O
    forced = _forced_keep_from(res, rows, B)
    sub = [[x[rows] for x in a] for a in inputs_cpu]
    ora = O.forward(variant, model_cpu_sd, cfg, *sub, forced_keep=forced)
    return ora, forced


@pytest.mark.parametrize("sharpen", [True, False], ids=["sharpened", "plain"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_fp32_mode_matches_reference_golden(built_lib, variant, sharpen):
    res, cfg = _run(variant, "fp32", sharpen=sharpen)
    g = _golden(variant, sharpen)
    d_box = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max()
    d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
    print(f"{variant} fp32: boxes {d_box:.3e} (norm.)  score maps {d_map:.3e}")
    assert d_box <= 1e-4, d_box          # north star, fp32 mode: 1e-4 (boxes: 1e-4 normalised = 0.03 px)
    assert d_map <= 1e-4, d_map          # absolute, on raw corner logits that reach |12| on the sharpened set
    if variant == "asymmetric_shared_ce":
        B = 2
        keeps = [k.cpu().numpy().astype(np.int32) for k in res["ce_keep"]]      # [2B, keep], modality-major
        for j in range(3):
            for m in range(2):
                assert _ce_stage_matches(keeps, g, j, m, B), f"CE stage {j} modality {m}: kept indices differ"


@pytest.mark.parametrize("variant", VARIANTS)
def test_bf16_mode_north_star_tolerance(built_lib, variant):
    """North-star bound, on the weight set it names (the builders' random init): boxes <= 0.5 px and corner score
    maps <= 1e-2 abs against the fp32 reference."""
    res, cfg = _run(variant, "bf16", sharpen=False)
    g = _golden(variant, False)
    size = cfg.DATA.SEARCH.SIZE
    d_box_px = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * size
    d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
    print(f"{variant} bf16 (plain init): boxes {d_box_px:.3f} px  score maps {d_map:.3e}")
    assert d_box_px <= 0.5, d_box_px
    if variant != "asymmetric_shared_ce":
        assert d_map <= 1e-2, d_map
    else:
        _check_ce_under_forced_keep(res, cfg, sharpen=False, precision="bf16", tol_box_px=0.5, tol_map=1e-2)


def _check_ce_under_forced_keep(res, cfg, sharpen, precision, tol_box_px, tol_map, batch=2, yaml_name=None):
    """asymmetric_shared_ce off the fp32 path: bf16 q/k move scores that are closer than bf16 resolution across the keep
    boundary, so a different (equally valid) token is pruned and single logits of the recovered map move.  The test
    separates the two effects: (1) the oracle is re-run with the GPU's kept sets FORCED (same token population at
    every stage) and the arithmetic bound is asserted against that run; (2) every token on which the GPU's choice
    differs from the oracle's own top-k must be a near-tie in the oracle's scores (CE_TIE_REL)."""
    from mmt_b200 import synthetic
    variant = "asymmetric_shared_ce"
    model, _ = synthetic.make_model(variant, 0, sharpen=sharpen, yaml_name=yaml_name)
    inputs = synthetic.make_inputs(variant, cfg, batch, 1)
    rows = torch.arange(batch)
    ora, forced = _ce_forced_oracle(variant, model.state_dict(), cfg, inputs, res, rows, batch)
    n_diff, worst = _assert_ce_keep_sets_are_score_consistent(ora, forced, CE_TIE_REL[precision])
    d_box_px = (res["pred_boxes"].cpu() - ora["pred_boxes"]).abs().max().item() * cfg.DATA.SEARCH.SIZE
    d_map = (res["score_maps"].cpu() - ora["score_maps"]).abs().max().item()
    print(f"   forced-keep oracle: boxes {d_box_px:.3f} px  score maps {d_map:.3e}; {n_diff} kept tokens differ from the "
          f"oracle's own top-k, worst relative distance from the boundary {worst:.2e}")
    assert d_box_px <= tol_box_px, d_box_px
    assert d_map <= tol_map, d_map


@pytest.mark.parametrize("variant", VARIANTS)
def test_bf16_mode_sharpened_weights(built_lib, variant):
    """Stress set: head gain x24 makes the corner logits reach |12| and amplifies every bf16 rounding.  The
    reference itself under torch.autocast(bfloat16) is 1.5 px / 0.17 away from its fp32 self on this set
    (DESIGN.md 'bf16 error budget'), so the bound here is 2 px and 2% of the logit range, not the north-star one.
    CE in bf16 may legitimately keep a different token set (near-tied scores): that variant is compared with the oracle
    in forced-keep mode and every differing token must be a near-tie (_check_ce_under_forced_keep)."""
    res, cfg = _run(variant, "bf16", sharpen=True)
    g = _golden(variant, True)
    size = cfg.DATA.SEARCH.SIZE
    d_box_px = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * size
    d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
    print(f"{variant} bf16 (sharpened): boxes {d_box_px:.3f} px  score maps {d_map:.3e}")
    if variant != "asymmetric_shared_ce":
        assert d_box_px <= 2.0, d_box_px
        assert d_map <= 2e-2 * np.abs(g["score_maps"]).max(), d_map
    else:
        _check_ce_under_forced_keep(res, cfg, sharpen=True, precision="bf16", tol_box_px=2.0,
                                    tol_map=2e-2 * float(np.abs(g["score_maps"]).max()))


def test_live_oracle_ragged_batch(built_lib):
    """Odd batch size (3) against the oracle run live on the CPU (no fixture for this shape)."""
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    variant = "mixformer_vit_rgbt_shared"
    model, cfg = synthetic.make_model(variant, 3)
    inputs = synthetic.make_inputs(variant, cfg, 3, 7)
    ora = O.forward(variant, model.state_dict(), cfg, *inputs)
    model = model.cuda().set_precision("fp32")
    cu = [[t.cuda() for t in x] for x in inputs]
    out, _ = model(*cu)
    d = (out["pred_boxes"].cpu() - ora["pred_boxes"]).abs().max().item()
    assert d <= 1e-4, d


def test_cpu_inputs_rejected(built_lib):
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model("mixformer_vit", 0)
    inputs = synthetic.make_inputs("mixformer_vit", cfg, 1, 1)
    with pytest.raises(NotImplementedError):
        model(*inputs)                      # model still on the CPU: no fallback
    model = model.cuda()
    with pytest.raises(NotImplementedError):
        model(*inputs)                      # CPU tensors into a CUDA model


def test_cuda_graph_replay_matches_eager(built_lib):
    """enable_cuda_graph: same boxes as the eager launch sequence, for changing inputs and two batch sizes."""
    from mmt_b200 import synthetic
    for variant in ("mixformer_vit_rgbt_shared", "asymmetric_shared_ce"):
        model, cfg = synthetic.make_model(variant, 0)
        model = model.cuda()
        for batch, seed in ((1, 3), (2, 4), (1, 5)):
            inputs = synthetic.make_inputs(variant, cfg, batch, seed, device="cuda")
            model.enable_cuda_graph(False)
            _, eager = model(*inputs)
            model.enable_cuda_graph(True)
            _, graphed = model(*inputs)
            _, graphed2 = model(*inputs)
            torch.cuda.synchronize()
            assert torch.equal(eager, graphed) and torch.equal(graphed, graphed2), (variant, batch)


def test_in_place_parameter_edits_repack_the_weights(built_lib):
    """In-place parameter edits after the first forward (which the load_state_dict / .cuda() hooks cannot see) reach the
    packed device arena - and the captured CUDA graphs - through invalidate()."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model("mixformer_vit", 0)
    model = model.cuda()
    inputs = synthetic.make_inputs("mixformer_vit", cfg, 2, 3, device="cuda")
    _, a = model(*inputs)
    with torch.no_grad():
        model.box_head.conv5_tl.bias.add_(3.0)                      # invisible to load_state_dict / _apply hooks
        model.box_head.conv5_tl.weight.mul_(1.5)
    model(*inputs)               # before invalidate(): unspecified (fp32 parameters the kernels read in place are seen,
                                 # packed / bf16 / folded copies are not) - only the state after invalidate() is a contract
    _, b = model.invalidate()(*inputs)
    fresh, _ = synthetic.make_model("mixformer_vit", 0)
    fresh.load_state_dict(model.state_dict())
    _, want = fresh.cuda()(*inputs)
    torch.cuda.synchronize()
    assert not torch.equal(a, b) and torch.equal(b, want)
    model.enable_cuda_graph(True)
    _, g1 = model(*inputs)
    with torch.no_grad():
        model.box_head.conv5_tl.weight.mul_(1.0 / 1.5)
        model.box_head.conv5_tl.bias.sub_(3.0)
    _, g2 = model.invalidate()(*inputs)
    torch.cuda.synchronize()
    assert torch.equal(g1, want) and torch.allclose(g2, a, atol=1e-6)


def test_single_sequence_sm_budget_is_result_neutral(built_lib):
    """engine._run_two_backbones gives each modality chain half the SMs at small batch (mmt_config_small_gemm_sms): tile widths
    change, results do not - eager and graph replay."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model("mixformer_vit_rgbt", 0)
    model = model.cuda()
    for batch in (1, 2):
        inputs = synthetic.make_inputs("mixformer_vit_rgbt", cfg, batch, 4, device="cuda")
        _, ref = model(*inputs)
        eng = model.engine()
        assert eng._half_sm_small
        eng._half_sm_small = False
        try:
            _, whole = model(*inputs)
        finally:
            eng._half_sm_small = True
        torch.cuda.synchronize()
        assert torch.equal(ref, whole), batch


def test_programmatic_dependent_launch_is_result_neutral(built_lib):
    """mmt_config_pdl: the kernels' prologues overlapping the previous kernel's tail (griddepcontrol) change no result -
    eager launches and graph replay, two variants (two streams / one stream with candidate elimination)."""
    from mmt_b200 import ops, synthetic
    prev = ops.config_pdl(True)
    try:
        for variant in ("mixformer_vit_rgbt", "asymmetric_shared_ce"):
            model, cfg = synthetic.make_model(variant, 0)
            model = model.cuda()
            for batch in (1, 3):
                inputs = synthetic.make_inputs(variant, cfg, batch, 7, device="cuda")
                ops.config_pdl(False)
                _, plain = model(*inputs)
                ops.config_pdl(True)
                _, pdl = model(*inputs)
                model.enable_cuda_graph(True)
                _, graphed = model(*inputs)
                _, graphed2 = model(*inputs)
                model.enable_cuda_graph(False)
                torch.cuda.synchronize()
                assert torch.equal(plain, pdl) and torch.equal(plain, graphed) and torch.equal(plain, graphed2), (variant, batch)
    finally:
        ops.config_pdl(prev)


ONLINE = "mixformer_vit_online"


ONLINE_VARIANTS = ["mixformer_vit_online", "mixformer_convmae_online"]


@pytest.mark.parametrize("sharpen", [True, False], ids=["sharpened", "plain"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ONLINE", ONLINE_VARIANTS)
def test_online_score_model(built_lib, ONLINE, precision, sharpen):
    """SPM score head (PrRoIPool + score-token decoder) on the full forward, and the cached-template path
    set_online + forward_test, against the reference's golden outputs (lib/models/mixformer_vit/mixformer_online.py)."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(ONLINE, 0, sharpen=sharpen)
    model = model.cuda().set_precision(precision)
    g = _golden(ONLINE, sharpen)
    inputs = synthetic.make_inputs(ONLINE, cfg, 2, 1, device="cuda")
    out, coords = model(*inputs, run_score_head=True)
    res = model._engine.forward(*inputs)
    torch.cuda.synchronize()
    assert torch.equal(res["pred_boxes"], coords) and torch.equal(res["pred_scores"], out["pred_scores"])
    tt, oo, ss = synthetic.make_online_inputs(cfg, 3, 11, device="cuda")
    model.set_online(tt, oo)
    out2, _ = model.forward_test(ss, run_score_head=True)
    res2 = model._engine.forward_test(ss)
    torch.cuda.synchronize()
    size = cfg.DATA.SEARCH.SIZE
    d = dict(box=np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * size,
             map=np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max(),
             score=np.abs(out["pred_scores"].cpu().numpy() - g["pred_scores"]).max(),
             obox=np.abs(out2["pred_boxes"].cpu().numpy() - g["online_pred_boxes"]).max() * size,
             omap=np.abs(res2["score_maps"].cpu().numpy() - g["online_score_maps"]).max(),
             oscore=np.abs(out2["pred_scores"].cpu().numpy() - g["online_pred_scores"]).max())
    print(f"{ONLINE} {precision} {'sharpened' if sharpen else 'plain'}: " + "  ".join(f"{k} {v:.3e}" for k, v in d.items()))
    assert out["pred_scores"].shape == (2,) and out2["pred_scores"].shape == (1,)
    if precision == "fp32":
        assert d["box"] <= 1e-4 * size and d["obox"] <= 1e-4 * size
        assert max(d["map"], d["omap"], d["score"], d["oscore"]) <= 1e-4
    elif not sharpen:        # north-star bf16 bound on the weight set it names; score logits to 1e-2 as well
        assert d["box"] <= 0.5 and d["obox"] <= 0.5
        assert max(d["map"], d["omap"], d["score"], d["oscore"]) <= 1e-2
    else:
        lim = 2e-2 * np.abs(g["score_maps"]).max()
        assert d["box"] <= 2.0 and d["obox"] <= 2.0 and d["map"] <= lim and d["omap"] <= lim
        assert max(d["score"], d["oscore"]) <= 5e-2


def test_online_paths_reject_misuse(built_lib):
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(ONLINE, 0)
    model = model.cuda()
    t, ot, s = synthetic.make_inputs(ONLINE, cfg, 2, 1, device="cuda")
    with pytest.raises(RuntimeError):
        model.forward_test(s[:1])                 # no cache yet
    with pytest.raises(RuntimeError):
        model.set_online(t, ot)                   # template batch must be 1
    model.set_online(t[:1], ot)
    with pytest.raises(RuntimeError):
        model.forward_test(s)                     # one search crop per cached sequence


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("ONLINE", ONLINE_VARIANTS)
def test_online_score_model_large(built_lib, ONLINE, precision):
    """The -L online models of BASELINE.json configs[4] (experiments/mixformer_{vit,convmae}_online/baseline_large.yaml:
    MixViT-L 1024 x 24 layers / ConvMAE-L 384-768-1024 stem + 20 layers, 16 heads, 384 search / 192 template -> 864
    tokens, 96 x 96 corner maps): full forward + SPM and the cached-template path."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(ONLINE, 0, sharpen=False, yaml_name="baseline_large")
    model = model.cuda().set_precision(precision)
    g = np.load(os.path.join(GOLDEN, f"{ONLINE}_plain_baseline_large_b1.npz"))
    inputs = synthetic.make_inputs(ONLINE, cfg, 1, 1, device="cuda")
    out, _ = model(*inputs, run_score_head=True)
    tt, oo, ss = synthetic.make_online_inputs(cfg, 3, 11, device="cuda")
    model.set_online(tt, oo)
    out2, _ = model.forward_test(ss, run_score_head=True)
    torch.cuda.synchronize()
    size = cfg.DATA.SEARCH.SIZE
    d = dict(box=np.abs(out["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * size,
             score=np.abs(out["pred_scores"].cpu().numpy() - g["pred_scores"]).max(),
             obox=np.abs(out2["pred_boxes"].cpu().numpy() - g["online_pred_boxes"]).max() * size,
             oscore=np.abs(out2["pred_scores"].cpu().numpy() - g["online_pred_scores"]).max())
    print(f"{ONLINE} large {precision}: " + "  ".join(f"{k} {v:.3e}" for k, v in d.items()))
    if precision == "fp32":
        assert max(d["box"], d["obox"]) <= 1e-4 * size and max(d["score"], d["oscore"]) <= 1e-4
    else:
        assert max(d["box"], d["obox"]) <= 0.5 and max(d["score"], d["oscore"]) <= 1e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("variant", ["mixformer_vit", "mixformer_vit_rgbt", "mixformer_vit_rgbt_shared",
                                     "mixformer_vit_rgbt_unibackbone", "asymmetric_shared", "asymmetric_shared_ce"])
def test_template_cache_is_bit_identical(built_lib, variant, precision):
    """cache_templates() + forward_search() (search tokens only, cached template q/k/v as a second key source) must
    reproduce forward() bit for bit: idempotence of the template side.  Cross-modal variants: the
    search rows read the cached template rows of BOTH modalities; candidate elimination: the scores' template queries
    come from the cache (mmt_ce_scores_split) and the kept sets must be identical too."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(variant, 0)
    model = model.cuda().set_precision(precision)
    t, ot, s = synthetic.make_inputs(variant, cfg, 3, 5, device="cuda")
    _, full = model(t, ot, s)
    model.cache_templates(t, ot)
    _, cached = model.forward_search(s)
    t2, ot2, s2 = synthetic.make_inputs(variant, cfg, 3, 6, device="cuda")      # new frame, same templates
    _, full2 = model(t, ot, s2)
    _, cached2 = model.forward_search(s2)
    torch.cuda.synchronize()
    assert torch.equal(full, cached) and torch.equal(full2, cached2)
    if variant == "asymmetric_shared_ce":
        eng = model.engine()
        k_full = [k.clone() for k in eng.forward(t, ot, s2)["ce_keep"]]
        k_cached = [k.clone() for k in eng.forward_search(s2)["ce_keep"]]
        assert len(k_full) == 3 and all(torch.equal(a, b) for a, b in zip(k_full, k_cached))
    with pytest.raises(RuntimeError):
        model.forward_search([x[:2] for x in s] if isinstance(s, list) else s[:2])      # batch differs from the cache


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("variant,yaml_name", [("mixformer_vit_rgbt_shared", "baseline_attention_lasher_newfusion_2layer"),
                                               ("asymmetric_shared", "attention_lasher_cat_3layer"),
                                               ("mixformer_vit", "baseline_large"),
                                               ("asymmetric_shared", "attention-lasher-cross_deform_fusion_sum_2layer"),
                                               ("asymmetric_shared", "attention_lasher_newfusionAdd_2layer"),
                                               ("asymmetric_shared_ce", "attention_lasher_newfusionAdd_2layer")])
def test_other_fusion_classes(built_lib, variant, yaml_name, precision):
    """The remaining fusion classes of the shipped YAMLs: Attention_Fusion_Bimodal (one LayerNorm for both modalities,
    deformable_encoder.py:111-158) and RGBT_Fusion_Cat (3 x conv3x3 + BN + ReLU on the channel concat,
    fusion_utils.py:86-110), against the reference's golden outputs; and MixViT-L RGB-only (experiments/mixformer_vit/
    baseline_large.yaml: 24 blocks x 1024, 384^2 search / 192^2 templates); Attention_Fusion_Bimodal_LNSpecific_Sum and
    _2 (fusion_utils.py:282-353: out_v + out_i through ONE 1x1 conv + GroupNorm; _2 also shares the input adjust conv
    between the modalities), the latter with and without candidate elimination."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(variant, 0, yaml_name=yaml_name)
    model = model.cuda().set_precision(precision)
    inputs = synthetic.make_inputs(variant, cfg, 2, 1, device="cuda")
    res = model.engine().forward(*inputs)
    torch.cuda.synchronize()
    g = np.load(os.path.join(GOLDEN, f"{variant}__{yaml_name}_b2.npz"))
    d_box = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * cfg.DATA.SEARCH.SIZE
    d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
    print(f"{variant}/{yaml_name} {cfg.MODEL.get('FUSION_CLASS')} {precision}: boxes {d_box:.3e} px  maps {d_map:.3e}")
    if precision == "fp32":
        assert d_box <= 1e-4 * cfg.DATA.SEARCH.SIZE and d_map <= 2e-4
    elif variant == "asymmetric_shared_ce":       # bf16 may flip near-tied tokens across the keep boundary
        # stress set: single logits of the recovered map moved by 1.9-2.3 % of the map maximum across the round's builds
        # (rounding-level changes in the attention softmax move it by that much): 2.5 %
        _check_ce_under_forced_keep(res, cfg, sharpen=True, precision="bf16", tol_box_px=2.0,
                                    tol_map=2.5e-2 * float(np.abs(g["score_maps"]).max()), yaml_name=yaml_name)
    else:
        # sharpened stress set (head gain x24): 2 px at the 288-px search crop, i.e. 6.9e-3 of the crop side - the bf16 error
        # lives in normalised coordinates, so the 384-px crops of the -L models get the same normalised bound (2.67 px);
        # profiles/r2_bf16_error_table.md has every variant with the LayerNorm fold on and off.
        # Where the reference's own map is MULTI-MODAL (the random -L weights put 0.37 / 0.29 of the soft-argmax mass on two
        # peaks 200 px apart) a fixed pixel bound tests the luck of the rounding, not the network.  Yardstick: the UNMODIFIED
        # reference in its own reduced-precision mode (torch.autocast(bfloat16), stored with the golden by oracle/gen_golden.py:
        # its boxes move 8.2 px and its maps 0.28 on this case) - the B200 path may not be worse than 1.25 x that.
        assert d_map <= 2e-2 * np.abs(g["score_maps"]).max()
        fixed = 2.0 * cfg.DATA.SEARCH.SIZE / 288.0
        ref_dev = (np.abs(g["pred_boxes_autocast_bf16"] - g["pred_boxes"]).max() * cfg.DATA.SEARCH.SIZE
                   if "pred_boxes_autocast_bf16" in g.files else 0.0)
        bound = max(fixed, 1.25 * ref_dev)
        print(f"  box bound {bound:.2f} px (fixed {fixed:.2f}, reference under bf16 autocast {ref_dev:.2f})")
        assert d_box <= bound


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_plain_corner_head(built_lib, precision):
    """HEAD_TYPE = CORNER (Corner_Predictor, lib/models/mixformer_cvt/head.py:23-94: conv tower at stride 16, 18 x 18 corner
    maps) behind the MixViT-B backbone, against the reference's outputs (oracle/gen_golden.py main_corner_head)."""
    from mmt_b200 import synthetic
    variant = "mixformer_vit"
    model, cfg = synthetic.make_model(variant, 0, overrides={"MODEL.HEAD_TYPE": "CORNER"})
    model = model.cuda().set_precision(precision)
    inputs = synthetic.make_inputs(variant, cfg, 2, 1, device="cuda")
    res = model.engine().forward(*inputs)
    torch.cuda.synchronize()
    g = np.load(os.path.join(GOLDEN, f"{variant}__head_corner_b2.npz"))
    assert res["score_maps"].shape == (2, 2, 18 * 18)
    d_box = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * cfg.DATA.SEARCH.SIZE
    d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
    print(f"{variant} HEAD_TYPE=CORNER {precision}: boxes {d_box:.3e} px  maps {d_map:.3e}")
    if precision == "fp32":
        assert d_box <= 1e-4 * cfg.DATA.SEARCH.SIZE and d_map <= 2e-4
    else:
        assert d_box <= 2.0 and d_map <= 2e-2 * np.abs(g["score_maps"]).max()


@pytest.mark.parametrize("variant,batch", [("mixformer_vit_rgbt", 64), ("asymmetric_shared_ce", 128)])
def test_full_size_batch_independence(built_lib, variant, batch):
    """BASELINE.json configs[1] / configs[3] at their full batch sizes, through size-independent properties: sequences
    are independent, so (a) a sub-batch gives bit-identical boxes, (b) permuting the sequences permutes the boxes,
    (c) every box is finite and inside the unit square (cx, cy of a soft-argmax) - in the bf16 tensor-core path."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(variant, 0)
    model = model.cuda()
    t, ot, s = synthetic.make_inputs(variant, cfg, batch, 3, device="cuda")
    _, full = model(t, ot, s)
    sub = lambda a, idx: [x[idx].contiguous() for x in a]
    _, first = model(sub(t, slice(0, 5)), sub(ot, slice(0, 5)), sub(s, slice(0, 5)))
    perm = torch.randperm(batch, generator=torch.Generator().manual_seed(0)).cuda()
    _, shuffled = model(sub(t, perm), sub(ot, perm), sub(s, perm))
    torch.cuda.synchronize()
    assert torch.equal(full[:5], first), "a sub-batch must reproduce its sequences bit for bit"
    assert torch.equal(full[perm], shuffled), "permuting the sequences must permute the boxes"
    b = full.view(-1, 4)
    assert bool(torch.isfinite(b).all()) and bool(((b[:, :2] >= 0) & (b[:, :2] <= 1)).all())
    assert b[:, :2].std().item() > 1e-3            # boxes differ between sequences (not a constant output)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("variant,batch", [("mixformer_vit_rgbt", 64), ("asymmetric_shared_ce", 128)])
def test_full_size_batch_against_live_oracle(built_lib, variant, batch, precision):
    """BASELINE.json configs[1] (two-stream, bs=64) and configs[3] (candidate elimination, bs=128) at their FULL batch
    sizes, on the weight set `north_star` names (the builders' random init): 8 sequences spread over the batch are
    re-computed by the CPU oracle (sequences are independent) and must meet the north-star bounds.  At these sizes the
    backbone GEMMs run in the cta_group::2 CTA-pair kernel and the attention kernel at full occupancy - the launch
    configurations bench.py times."""
    from mmt_b200 import synthetic
    from oracle import mixformer_oracle as O
    model, cfg = synthetic.make_model(variant, 0, sharpen=False)
    sd = model.state_dict()
    inputs = synthetic.make_inputs(variant, cfg, batch, 3)
    model = model.cuda().set_precision(precision)
    cu = [[x.cuda() for x in a] for a in inputs]
    res = model.engine().forward(*cu)
    torch.cuda.synchronize()
    rows = torch.tensor([0, 1, batch // 4 + 3, batch // 2 - 1, batch // 2, 3 * batch // 4 + 2, batch - 2, batch - 1])
    size = cfg.DATA.SEARCH.SIZE
    tol_box, tol_map = (0.5 / size, 1e-2) if precision == "bf16" else (1e-4, 1e-4)
    if variant == "asymmetric_shared_ce":
        ora, forced = _ce_forced_oracle(variant, sd, cfg, inputs, res, rows, batch)
        n_diff, worst = _assert_ce_keep_sets_are_score_consistent(ora, forced, CE_TIE_REL[precision])
        print(f"   CE: {n_diff} kept tokens differ from the oracle's own top-k (worst relative gap {worst:.2e})")
        if precision == "fp32":
            assert n_diff == 0 or worst <= CE_TIE_REL["fp32"]
    else:
        ora = O.forward(variant, sd, cfg, *[[x[rows] for x in a] for a in inputs])
    d_box = (res["pred_boxes"].cpu()[rows] - ora["pred_boxes"]).abs().max().item()
    d_map = (res["score_maps"].cpu()[rows] - ora["score_maps"]).abs().max().item()
    print(f"{variant} bs={batch} {precision}: 8 sequences vs live oracle: boxes {d_box * size:.4f} px  maps {d_map:.3e}")
    assert d_box <= tol_box, d_box * size
    assert d_map <= tol_map, d_map


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("weights", ["", "_plain"])
def test_asymmetric_shared_online(built_lib, weights, precision):
    """asymmetric_shared + SPM score head on the fused map (lib/models/mixformer_vit_rgbt/asymmetric_shared_online.py:
    337-413, run_score_head=True) against the reference's golden boxes, corner maps and score logits."""
    from mmt_b200 import synthetic
    variant = "asymmetric_shared_online"
    model, cfg = synthetic.make_model(variant, 0, sharpen=(weights == ""))
    model = model.cuda().set_precision(precision)
    inputs = synthetic.make_inputs(variant, cfg, 2, 1, device="cuda")
    res = model.engine().forward(*inputs, run_score_head=True)
    out, coords = model(*inputs, run_score_head=True)
    torch.cuda.synchronize()
    g = np.load(os.path.join(GOLDEN, f"{variant}__spm{weights}_b2.npz"))
    d_box = np.abs(res["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max() * cfg.DATA.SEARCH.SIZE
    d_map = np.abs(res["score_maps"].cpu().numpy() - g["score_maps"]).max()
    d_sc = np.abs(out["pred_scores"].cpu().numpy() - g["pred_scores"]).max()
    print(f"{variant}{weights} {precision}: boxes {d_box:.3e} px  maps {d_map:.3e}  scores {d_sc:.3e}")
    assert torch.equal(out["pred_boxes"], res["pred_boxes"]) and out["pred_scores"].shape == (2,)
    if precision == "fp32":
        assert d_box <= 1e-4 * cfg.DATA.SEARCH.SIZE and d_map <= 2e-4 and d_sc <= 1e-4
    elif weights == "_plain":
        assert d_box <= 0.5 and d_map <= 1e-2 and d_sc <= 1e-2            # north-star bf16 bounds on random-init weights
    else:
        assert d_box <= 2.0 and d_map <= 2e-2 * np.abs(g["score_maps"]).max() and d_sc <= 2e-2
    # without the score head the reference returns boxes only (run_score_head defaults to False, :352)
    out2, _ = model(*inputs)
    assert "pred_scores" not in out2 and torch.equal(out2["pred_boxes"], out["pred_boxes"])
    with pytest.raises(NotImplementedError):
        model.engine().set_online(None, None)


@pytest.mark.parametrize("variant,B,n", [("mixformer_vit_online", 3, 2), ("mixformer_convmae_online", 2, 3)])
def test_batched_cached_template_path(built_lib, variant, B, n):
    """set_online_batch / forward_test_batch (B sequences, n online templates each) against the reference-shaped batch-1
    set_online / forward_test run per sequence: boxes, corner maps and score logits bit-identical."""
    from mmt_b200 import synthetic
    model, cfg = synthetic.make_model(variant, 0)
    model = model.cuda()
    g = torch.Generator().manual_seed(5)
    ts, ss = cfg.DATA.TEMPLATE.SIZE, cfg.DATA.SEARCH.SIZE
    t = torch.randn(B, 3, ts, ts, generator=g).cuda()
    ot = torch.randn(B, n, 3, ts, ts, generator=g).cuda()
    s = torch.randn(B, 3, ss, ss, generator=g).cuda()
    model.set_online_batch(t, ot)
    out, coords = model.forward_test_batch(s, run_score_head=True)
    maps = model.engine().forward_test_batch(s)["score_maps"].clone()
    assert coords.shape == (B, 1, 4) and out["pred_scores"].shape == (B,)
    for b in range(B):
        model.set_online(t[b:b + 1], ot[b])
        o1, c1 = model.forward_test(s[b:b + 1], run_score_head=True)
        m1 = model.engine().forward_test(s[b:b + 1])["score_maps"]
        assert torch.equal(c1[0], coords[b]), (b, (c1[0] - coords[b]).abs().max().item())
        assert torch.equal(o1["pred_scores"][0], out["pred_scores"][b])
        assert torch.equal(m1[0], maps[b])
    with pytest.raises(RuntimeError):
        model.forward_test_batch(s[:1])
