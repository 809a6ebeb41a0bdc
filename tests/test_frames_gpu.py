"""GPU parity of the device-side frame path (csrc/frames.cu through the C ABI, mmt_b200/frames.py) against
oracle/frame_oracle.py and the fixture produced by the UNMODIFIED reference functions (tests/golden/frames_rgbt.npz):
uint8 crops bit-exact, normalised fp32 crops bit-exact, float64 tracker state exactly equal, and a closed-loop
multi-frame run of BatchedTracker against the reference's per-sequence host loop restated with the oracle."""
import types

import numpy as np
import pytest
import torch

from oracle import frame_oracle as FO
from oracle import gen_golden_frames as GG

pytestmark = pytest.mark.gpu

GOLD = np.load(GG.os.path.join(GG.GOLDEN, "frames_rgbt.npz"))
N_CASES = int(GOLD["n_cases"])
T_FACTOR, T_SIZE, S_FACTOR, S_SIZE = [float(v) for v in GOLD["params"]]


def _upload(images):
    """list of uint8 HWC arrays -> (device buffers kept alive, int64 pointer table, int32 dims table)."""
    bufs = [torch.from_numpy(np.ascontiguousarray(im)).cuda() for im in images]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    dims = torch.tensor([[im.shape[0], im.shape[1], im.shape[1] * 3] for im in images], dtype=torch.int32, device="cuda")
    return bufs, ptrs, dims


def test_crops_bit_exact_against_reference_fixture(built_lib):
    """All fixture cases as ONE batch (B = 8 sequences with different frame sizes, 2 modalities)."""
    from mmt_b200 import ops, frames
    vis, inf = zip(*[GG.case_frames(ci) for ci in range(N_CASES)])
    bufs, ptrs, dims = _upload(list(vis) + list(inf))
    state = torch.tensor(np.stack([GOLD[f"c{ci}_box"] for ci in range(N_CASES)]), device="cuda")
    lut = frames.jet_lut_tensor("cuda")
    for name, factor, size in (("template", T_FACTOR, int(T_SIZE)), ("search", S_FACTOR, int(S_SIZE))):
        u8 = torch.empty((2, N_CASES, size, size, 3), dtype=torch.uint8, device="cuda")
        out, _, rf = ops.frame_crop(ptrs, dims, state, factor, size, 2, jet_mask=0b10, jet_lut=lut, out_u8=u8,
                                    out=torch.empty((2, N_CASES, 3, size, size), device="cuda"))
        torch.cuda.synchronize()
        u8, out, rf = u8.cpu().numpy(), out.cpu().numpy(), rf.cpu().numpy()
        for ci in range(N_CASES):
            assert rf[ci] == float(GOLD[f"c{ci}_{name}_rf"])
            for m, tag in ((0, "v"), (1, "i")):
                assert GG.sha(u8[m, ci]) == str(GOLD[f"c{ci}_{name}_u8_{tag}_sha"]), (ci, name, tag, "uint8 crop")
                assert GG.sha(out[m, ci]) == str(GOLD[f"c{ci}_{name}_{tag}_sha"]), (ci, name, tag, "normalised crop")
            if ci in GG.FULL_CASES:
                assert np.array_equal(u8[0, ci], GOLD[f"c{ci}_{name}_u8_v"])
                assert np.array_equal(u8[1, ci], GOLD[f"c{ci}_{name}_u8_i"])


def test_crop_random_boxes_match_oracle(built_lib):
    """Seeded boxes all over (and partly outside) frames of odd sizes, several output sizes; active mask honoured."""
    from mmt_b200 import ops
    rng = np.random.default_rng(3)
    B = 12
    images, boxes = [], []
    for b in range(B):
        H, W = int(rng.integers(40, 300)), int(rng.integers(40, 400))
        images.append(rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
        w, h = rng.uniform(8, W), rng.uniform(8, H)
        boxes.append([rng.uniform(-0.3 * w, W - 0.7 * w), rng.uniform(-0.3 * h, H - 0.7 * h), w, h])
    bufs, ptrs, dims = _upload(images)
    state = torch.tensor(np.array(boxes, dtype=np.float64), device="cuda")
    for factor, size in ((2.0, 128), (4.5, 288), (5.0, 320), (1.3, 37)):
        u8 = torch.full((1, B, size, size, 3), 7, dtype=torch.uint8, device="cuda")
        active = torch.ones(B, dtype=torch.uint8, device="cuda")
        active[3] = 0
        _, _, rf = ops.frame_crop(ptrs, dims, state, factor, size, 1, out_u8=u8, active=active)
        got = u8.cpu().numpy()[0]
        for b in range(B):
            if b == 3:
                assert (got[b] == 7).all()                     # inactive sequence untouched
                continue
            want, wrf = FO.sample_target(images[b], boxes[b], factor, size)
            assert np.array_equal(got[b], want), (b, factor, size)
            assert rf[b].item() == wrf


def test_track_update_matches_reference_fixture(built_lib):
    from mmt_b200 import ops
    for ci in range(N_CASES):
        H, W = GG.CASES[ci][0], GG.CASES[ci][1]
        preds = GOLD[f"c{ci}_pred_boxes"]
        n = preds.shape[0]
        state = torch.tensor(np.tile(GOLD[f"c{ci}_box"], (n, 1)), device="cuda")
        rf = torch.full((n,), float(GOLD[f"c{ci}_search_rf"]), dtype=torch.float64, device="cuda")
        dims = torch.tensor([[H, W, W * 3]] * n, dtype=torch.int32, device="cuda")
        log = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
        active = torch.ones(n, dtype=torch.uint8, device="cuda")
        active[n - 1] = 0
        ops.track_update(torch.from_numpy(preds).cuda(), rf, dims, state, int(S_SIZE), 10.0, log=log, active=active)
        got = state.cpu().numpy()
        assert np.array_equal(got[:n - 1], GOLD[f"c{ci}_next_states"][:n - 1])     # float64, exactly
        assert np.array_equal(got[n - 1], GOLD[f"c{ci}_box"])                       # inactive: state kept
        assert np.array_equal(log.cpu().numpy(), got)


def _video(rng, H, W, T):
    """A drifting textured scene: every frame differs, so stale crops would be caught."""
    base = rng.integers(0, 256, (H + 2 * T, W + 2 * T, 3), dtype=np.uint8)
    return [np.ascontiguousarray(base[t:t + H, 2 * t // 2:2 * t // 2 + W]) for t in range(T)]


@pytest.mark.parametrize("variant,use_cache", [("mixformer_vit_rgbt_shared", False), ("mixformer_vit_rgbt_shared", True),
                                               ("asymmetric_shared_ce", False), ("mixformer_vit", False)])
def test_batched_tracker_closed_loop_matches_host_loop(built_lib, variant, use_cache):
    """6 frames x 3 sequences (different frame sizes) with an online-template update at frame 4: BatchedTracker's
    float64 states must EQUAL those of the reference's per-sequence loop (asymmetric_shared_ce.py:74-111) restated
    with the oracle crops and the same network, frame after frame."""
    from mmt_b200 import synthetic, frames
    model, cfg = synthetic.make_model(variant, 0, sharpen=True)
    model = model.cuda()
    n_mod = 1 if variant == "mixformer_vit" else 2
    params = types.SimpleNamespace(template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE)
    rng = np.random.default_rng(11)
    sizes = [(240, 320), (200, 260), (301, 333)]
    T, B = 6, len(sizes)
    vids = [[_video(rng, H, W, T) for _ in range(n_mod)] for (H, W) in sizes]          # [b][m][t]
    init = np.array([[100.0, 80.0, 60.0, 50.0], [20.5, 30.25, 90.0, 40.0], [150.0, 150.0, 33.0, 71.0]])
    frames_at = lambda t: [[vids[b][m][t] for m in range(n_mod)] if n_mod > 1 else vids[b][0][t] for b in range(B)]

    trk = frames.BatchedTracker(model, params, update_intervals=[4], n_mod=n_mod, use_template_cache=use_cache)
    trk.initialize(frames_at(0), init)
    for t in range(1, T):
        trk.track(frames_at(t))
    got = trk.results()
    assert got.shape == (T, B, 4)

    # host loop, sequence by sequence through the oracle; the network itself is the same CUDA model, called on the
    # batch of oracle crops
    def crops(t, states, factor, size):
        per_mod = []
        for m in range(n_mod):
            arr = []
            for b in range(B):
                c, rf = FO.sample_target(vids[b][m][t], states[b], factor, size)
                arr.append(FO.normalize(FO.apply_jet(c) if m == 1 else c))
            per_mod.append(torch.from_numpy(np.stack(arr)).cuda())
        rfs = [size / FO.crop_geometry(states[b], factor, *sizes[b])[0] for b in range(B)]
        return (per_mod if n_mod > 1 else per_mod[0]), rfs

    states = [list(map(float, init[b])) for b in range(B)]
    template, _ = crops(0, states, params.template_factor, params.template_size)
    online = template
    want = [np.array(states)]
    for t in range(1, T):
        search, rfs = crops(t, states, params.search_factor, params.search_size)
        _, coords = model(template, online, search)
        pred = coords.view(-1, 4).cpu().numpy()
        states = [FO.update_state(states[b], pred[b], rfs[b], params.search_size, *sizes[b], margin=10) for b in range(B)]
        if t % 4 == 0:
            online, _ = crops(t, states, params.template_factor, params.template_size)
        want.append(np.array(states, dtype=np.float64))
    want = np.stack(want)
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert not np.array_equal(got[1], got[T - 1])              # the boxes actually moved


def test_run_sequences_refills_slots_and_writes_reference_files(built_lib, tmp_path):
    """5 sequences of different lengths / frame sizes through 2 slots: every sequence's boxes must equal the reference's
    one-sequence-at-a-time loop (oracle crops, same network, batch 1), whatever shared the batch with it; the result
    files have the reference's format (running.py:31-37: int boxes, tab separated; "%f" times)."""
    from mmt_b200 import synthetic, evaluation
    variant = "mixformer_vit_rgbt_shared"
    model, cfg = synthetic.make_model(variant, 0, sharpen=True)
    model = model.cuda()
    params = types.SimpleNamespace(template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE)
    rng = np.random.default_rng(21)
    shapes = [(120, 160, 4), (200, 150, 7), (90, 90, 1), (160, 240, 5), (131, 177, 6)]      # H, W, frames
    seqs = []
    for i, (H, W, T) in enumerate(shapes):
        vids = [_video(rng, H, W, T), _video(rng, H, W, T)]
        box = [W * 0.3, H * 0.25, W * 0.3, H * 0.4]
        seqs.append(evaluation.SequenceSpec(f"seq{i}", "synthetic", list(zip(vids[0], vids[1])), box))
    got = evaluation.run_sequences(model, params, seqs, results_dir=str(tmp_path), batch=2, update_intervals=[3], n_mod=2)
    assert sorted(got) == [f"seq{i}" for i in range(5)]

    for s in seqs:
        H, W = s.frames[0][0].shape[:2]
        state = list(s.init_bbox)

        def crop(t, state, factor, size):
            arr = []
            for m in range(2):
                c, _ = FO.sample_target(s.frames[t][m], state, factor, size)
                arr.append(torch.from_numpy(FO.normalize(FO.apply_jet(c) if m == 1 else c)[None]).cuda())
            return arr, size / FO.crop_geometry(state, factor, H, W)[0]

        template, _ = crop(0, state, params.template_factor, params.template_size)
        online = template
        want = [list(state)]
        for t in range(1, len(s.frames)):
            search, rf = crop(t, state, params.search_factor, params.search_size)
            _, coords = model(template, online, search)
            state = FO.update_state(state, coords.view(-1, 4).cpu().numpy()[0], rf, params.search_size, H, W, margin=10)
            if t % 3 == 0:
                online, _ = crop(t, state, params.template_factor, params.template_size)
            want.append([float(v) for v in state])
        want = np.array(want, dtype=np.float64)
        assert got[s.name].shape == want.shape
        assert np.array_equal(got[s.name], want), (s.name, np.abs(got[s.name] - want).max())
        saved = np.loadtxt(tmp_path / "synthetic" / f"{s.name}.txt", delimiter="\t", ndmin=2)
        assert np.array_equal(saved, want.astype(int))
        times = np.loadtxt(tmp_path / "synthetic" / f"{s.name}_time.txt", ndmin=1)
        assert times.shape == (len(s.frames),) and (times >= 0).all()


class _ScriptedOnlineNet(torch.nn.Module):
    """Stands in for the online model: replays scripted (box, logit) pairs per sequence and records digests of the
    crops it is given - so that the device-side bookkeeping can be compared with the reference-run fixture."""

    def __init__(self, scripts):
        super().__init__()
        self.anchor = torch.nn.Parameter(torch.zeros(1))
        self.scripts = scripts          # per sequence: (preds [T,4] f32, logits [T] f32)
        self.t = 0
        self.seen = []

    def forward(self, template, online_template, search, run_score_head=True):
        self.t += 1
        torch.cuda.synchronize()
        self.seen.append([(GG.sha(template[b].cpu().numpy()), GG.sha(online_template[b].cpu().numpy()),
                           GG.sha(search[b].cpu().numpy())) for b in range(search.shape[0])])
        boxes = torch.tensor(np.stack([p[self.t] for p, _ in self.scripts]), device="cuda").view(-1, 1, 4)
        logits = torch.tensor(np.array([l[self.t] for _, l in self.scripts], dtype=np.float32), device="cuda")
        return {"pred_boxes": boxes, "pred_scores": logits}, boxes


def test_online_batched_tracker_matches_reference_fixture(built_lib):
    """OnlineBatchedTracker (score-driven online-template candidate, periodic commit) with a scripted network: slots 0 and
    2 replay the fixture sequence that went through the UNMODIFIED MixFormerOnline class - their float64 states and the
    digests of every crop handed to the network must equal the fixture; slot 1 runs a different sequence / script."""
    from mmt_b200 import frames
    o = GG.ONLINE
    vid = GG.seeded_video(o["seed"], o["H"], o["W"], o["T"])
    other = GG.seeded_video(7, 150, 190, o["T"])
    preds, logits = GG.online_script(o["T"])
    preds2, logits2 = preds[::-1].copy(), (-logits[::-1]).copy()
    net = _ScriptedOnlineNet([(preds, logits), (preds2, logits2), (preds, logits)]).cuda()
    params = types.SimpleNamespace(template_factor=o["template_factor"], template_size=o["template_size"],
                                   search_factor=o["search_factor"], search_size=o["search_size"])
    trk = frames.OnlineBatchedTracker(net, params, update_interval=o["update_interval"])
    trk.initialize([vid[0], other[0], vid[0]], [o["box"], (40.0, 30.0, 50.0, 60.0), o["box"]])
    for t in range(1, o["T"]):
        trk.track([vid[t], other[t], vid[t]])
    got = trk.results()
    for slot in (0, 2):
        assert np.array_equal(got[:, slot], GOLD["online_states"])
        assert np.array_equal(np.array([s[slot] for s in net.seen]), GOLD["online_inputs_sha"])
    # slot 1 against the oracle loop with its own script
    seen = []

    def net_o(template, online_template, search):
        seen.append((GG.sha(template), GG.sha(online_template), GG.sha(search)))
        return preds2[orc.frame_id], logits2[orc.frame_id]

    orc = FO.OnlineTrackerOracle(net_o, o["template_factor"], o["template_size"], o["search_factor"], o["search_size"],
                                 o["update_interval"])
    orc.initialize(other[0], (40.0, 30.0, 50.0, 60.0))
    want = [[40.0, 30.0, 50.0, 60.0]] + [[float(v) for v in orc.track(other[t])] for t in range(1, o["T"])]
    assert np.array_equal(got[:, 1], np.array(want))
    assert [s[1] for s in net.seen] == seen
    assert np.array_equal(trk.score_logits()[1:, 0], logits[1:])


def test_online_batched_tracker_with_the_real_online_model(built_lib):
    """Closed loop with mixformer_vit_online (SPM score head on the device): states equal the per-sequence oracle loop
    driven by the same model."""
    from mmt_b200 import synthetic, frames
    model, cfg = synthetic.make_model("mixformer_vit_online", 0, sharpen=True)
    model = model.cuda()
    params = types.SimpleNamespace(template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE)
    T = 7
    vids = [GG.seeded_video(31, 180, 240, T), GG.seeded_video(32, 222, 201, T)]
    init = [(80.0, 60.0, 50.0, 40.0), (30.0, 90.0, 70.0, 45.0)]
    trk = frames.OnlineBatchedTracker(model, params, update_interval=3)
    trk.initialize([v[0] for v in vids], init)
    for t in range(1, T):
        trk.track([v[t] for v in vids])
    got, got_logits = trk.results(), trk.score_logits()
    for b in range(2):
        def net(template, online_template, search):
            out, coords = model(*[torch.from_numpy(a[None]).cuda() for a in (template, online_template, search)],
                                run_score_head=True)
            return coords.view(-1, 4).cpu().numpy()[0], out["pred_scores"].reshape(-1).cpu().numpy()[0]

        orc = FO.OnlineTrackerOracle(net, 2.0, params.template_size, 4.5, params.search_size, 3)
        orc.initialize(vids[b][0], init[b])
        want = [list(map(float, init[b]))] + [[float(v) for v in orc.track(vids[b][t])] for t in range(1, T)]
        assert np.array_equal(got[:, b], np.array(want)), (b, np.abs(got[:, b] - np.array(want)).max())
    assert np.isfinite(got_logits).all()


def test_preprocess_u8_matches_oracle(built_lib):
    """mmt_preprocess_u8 (Preprocessor_*.process on the device) bit for bit against the oracle, with and without JET."""
    from mmt_b200 import ops, frames
    rng = np.random.default_rng(17)
    crops = rng.integers(0, 256, (6, 37, 37, 3), dtype=np.uint8)
    lut = frames.jet_lut_tensor("cuda")
    out = torch.empty((6, 3, 37, 37), device="cuda")
    ops.preprocess_u8(torch.from_numpy(crops).cuda(), out, per_mod=3, jet_mask=0b10, jet_lut=lut)   # images 3..5: modality 1
    got = out.cpu().numpy()
    for i in range(6):
        want = FO.normalize(FO.apply_jet(crops[i]) if i >= 3 else crops[i])
        assert np.array_equal(got[i], want), i


@pytest.mark.parametrize("variant", ["mixformer_vit_rgbt_shared", "mixformer_vit", "asymmetric_shared_online",
                                     "mixformer_vit_rgbt", "asymmetric_shared_ce", "mixformer_vit_online"])
def test_framestep_uint8_crops_and_resident_templates(built_lib, variant):
    """FrameStep with uint8 HWC crops (uploaded as bytes, normalised on the device) gives the boxes of the model called
    on oracle-normalised fp32 crops; resident templates + step(None, None, search) gives the same boxes again."""
    from mmt_b200 import synthetic, runner
    model, cfg = synthetic.make_model(variant, 0, sharpen=True)
    model = model.cuda()
    rgbt = variant not in ("mixformer_vit", "mixformer_vit_online")
    B, ts, ss = 3, cfg.DATA.TEMPLATE.SIZE, cfg.DATA.SEARCH.SIZE
    rng = np.random.default_rng(23)
    u8 = lambda size: rng.integers(0, 256, (B, size, size, 3), dtype=np.uint8)
    norm = lambda a, m: torch.from_numpy(np.stack([FO.normalize(FO.apply_jet(x) if m == 1 else x) for x in a])).cuda()
    if rgbt:
        t, ot, s = [u8(ts), u8(ts)], [u8(ts), u8(ts)], [u8(ss), u8(ss)]
        ref_args = [[norm(a[m], m) for m in range(2)] for a in (t, ot, s)]
        host = [[torch.from_numpy(a[m]) for m in range(2)] for a in (t, ot, s)]
    else:
        t, ot, s = u8(ts), u8(ts), u8(ss)
        ref_args = [norm(a, 0) for a in (t, ot, s)]
        host = [torch.from_numpy(a) for a in (t, ot, s)]
    _, want = model(*ref_args)
    want = want.view(-1, 4).cpu()
    fs = runner.FrameStep(model)
    got = fs.step(*host).clone()
    assert torch.equal(got, want)
    assert fs.h2d_bytes == (2 if rgbt else 1) * B * 3 * (2 * ts * ts + ss * ss)          # one byte per value
    fs.set_templates(host[0], host[1])
    got2 = fs.step(None, None, host[2]).clone()
    assert torch.equal(got2, want)
    assert fs.h2d_bytes == (2 if rgbt else 1) * B * 3 * ss * ss
    with pytest.raises(RuntimeError):
        runner.FrameStep(model).step(None, None, host[2])


def test_dropin_tracker_class_protocol(built_lib):
    """`trackers.get_tracker_class(variant)(params, dataset)` with the reference's initialize / track protocol
    (lib/test/evaluation/tracker_rgbt.py:100-184): per-frame {"target_bbox": [x, y, w, h]} equal to the reference loop
    restated with the oracle and the same network; the online class against OnlineTrackerOracle."""
    from mmt_b200 import synthetic, trackers
    T = 6
    # RGB-T, Preprocessor_Multimodal variant
    cfg = synthetic.load_variant_config("mixformer_vit_rgbt_unibackbone")
    params = types.SimpleNamespace(cfg=cfg, template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE, checkpoint=None, save_all_boxes=False)
    torch.manual_seed(0)
    cls = trackers.get_tracker_class("mixformer_vit_rgbt_unibackbone")
    trk = cls(params, "lasher")
    synthetic.sharpen_(trk.network, torch.Generator().manual_seed(1000))
    trk.network.load_state_dict(trk.network.state_dict())                 # re-pack the engine arena after the edit
    trk.update_intervals = [2]
    vid = [GG.seeded_video(51, 190, 250, T), GG.seeded_video(52, 190, 250, T)]
    box = [70.0, 50.0, 66.0, 48.0]
    assert trk.initialize([vid[0][0], vid[1][0]], {"init_bbox": (box, box)}) is None
    got = [trk.track([vid[0][t], vid[1][t]])["target_bbox"] for t in range(1, T)]
    assert all(isinstance(g, list) and len(g) == 4 and all(isinstance(v, float) for v in g) for g in got)

    def crops(t, state, factor, size):
        out = []
        for m in range(2):
            c, _ = FO.sample_target(vid[m][t], state, factor, size)
            out.append(torch.from_numpy(FO.normalize(FO.apply_jet(c) if m == 1 else c)[None]).cuda())
        return out, size / FO.crop_geometry(state, factor, 190, 250)[0]

    state = list(box)
    template, _ = crops(0, state, 2.0, params.template_size)
    online = template
    for t in range(1, T):
        search, rf = crops(t, state, 4.5, params.search_size)
        _, coords = trk.network(template, online, search)
        state = FO.update_state(state, coords.view(-1, 4).cpu().numpy()[0], rf, params.search_size, 190, 250, margin=10)
        if t % 2 == 0:
            online, _ = crops(t, state, 2.0, params.template_size)
        assert got[t - 1] == [float(v) for v in state], t

    # online (SPM) class
    cfg = synthetic.load_variant_config("mixformer_vit_online")
    params = types.SimpleNamespace(cfg=cfg, template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE, checkpoint=None, save_all_boxes=False,
                                   update_interval=2)
    torch.manual_seed(0)
    otrk = trackers.get_tracker_class("mixformer_vit_online")(params, "lasot")
    otrk.initialize(vid[0][0], {"init_bbox": box})
    got = [otrk.track(vid[0][t])["target_bbox"] for t in range(1, T)]

    def net(template, online_template, search):
        out, coords = otrk.network(*[torch.from_numpy(a[None]).cuda() for a in (template, online_template, search)],
                                   run_score_head=True)
        return coords.view(-1, 4).cpu().numpy()[0], out["pred_scores"].reshape(-1).cpu().numpy()[0]

    orc = FO.OnlineTrackerOracle(net, 2.0, params.template_size, 4.5, params.search_size, 2)
    orc.initialize(vid[0][0], box)
    for t in range(1, T):
        assert got[t - 1] == [float(v) for v in orc.track(vid[0][t])], t
    with pytest.raises(KeyError):
        trackers.get_tracker_class("no_such_variant")


def test_batched_tracker_full_size_batch_independence(built_lib):
    """BASELINE.json size (64 sequences per GPU): three of the 64 sequences tracked alone (B = 3) give exactly the
    states they get inside the full batch - crops, forward (CTA-pair GEMMs at M = 57 856 vs single-CTA GEMMs at
    M = 2712, persistent attention over 3072 vs 144 items) and state update are all per-sequence computations."""
    from mmt_b200 import synthetic, frames
    variant = "mixformer_vit_rgbt_shared"
    model, cfg = synthetic.make_model(variant, 0, sharpen=True)
    model = model.cuda()
    params = types.SimpleNamespace(template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE)
    B, T, H, W = 64, 4, 120, 160
    rng = np.random.default_rng(77)
    vids = [[_video(rng, H, W, T), _video(rng, H, W, T)] for _ in range(B)]
    init = np.stack([rng.uniform(20, 80, B), rng.uniform(20, 60, B), rng.uniform(20, 60, B), rng.uniform(20, 50, B)], 1)
    pick = [0, 17, 63]

    def run(idx):
        trk = frames.BatchedTracker(model, params, update_intervals=[2], n_mod=2)
        trk.initialize([[vids[b][0][0], vids[b][1][0]] for b in idx], init[idx])
        for t in range(1, T):
            trk.track([[vids[b][0][t], vids[b][1][t]] for b in idx])
        return trk.results()

    full = run(list(range(B)))
    sub = run(pick)
    assert full.shape == (T, B, 4) and np.isfinite(full).all()
    assert np.array_equal(full[:, pick], sub)
    assert len({tuple(r) for r in full[T - 1].round(3).tolist()}) > B // 2       # the sequences really differ


def test_rgbt_online_tracker_class(built_lib):
    """`get_tracker_class("asymmetric_shared_online")` (lib/test/tracker/asymmetric_shared_online.py: RGB-T crops, SPM
    score on the fused map, score-driven online-template candidate) against the oracle loop with the same network."""
    from mmt_b200 import synthetic, trackers
    cfg = synthetic.load_variant_config("asymmetric_shared_online")
    params = types.SimpleNamespace(cfg=cfg, template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE, checkpoint=None, save_all_boxes=False,
                                   update_interval=2)
    torch.manual_seed(0)
    trk = trackers.get_tracker_class("asymmetric_shared_online")(params, "lasher")
    synthetic.sharpen_(trk.network, torch.Generator().manual_seed(1000))
    trk.network.load_state_dict(trk.network.state_dict())                 # re-pack the engine arena after the edit
    T, H, W = 6, 170, 230
    vid = [GG.seeded_video(61, H, W, T), GG.seeded_video(62, H, W, T)]
    box = [60.0, 45.0, 70.0, 52.0]
    trk.initialize([vid[0][0], vid[1][0]], {"init_bbox": (box, box)})
    got = [trk.track([vid[0][t], vid[1][t]])["target_bbox"] for t in range(1, T)]

    def net(template, online_template, search):
        to = lambda pair: [torch.from_numpy(a[None]).cuda() for a in pair]
        out, coords = trk.network(to(template), to(online_template), to(search), run_score_head=True)
        return coords.view(-1, 4).cpu().numpy()[0], out["pred_scores"].reshape(-1).cpu().numpy()[0]

    orc = FO.OnlineTrackerOracle(net, 2.0, params.template_size, 4.5, params.search_size, 2, rgbt=True)
    orc.initialize([vid[0][0], vid[1][0]], box)
    for t in range(1, T):
        assert got[t - 1] == [float(v) for v in orc.track([vid[0][t], vid[1][t]])], t


class _ScriptedCachedNet(torch.nn.Module):
    """Stub with the batched cached-template API: records what set_online_batch / forward_test_batch are given."""

    def __init__(self, preds, logits, B):
        super().__init__()
        self.anchor = torch.nn.Parameter(torch.zeros(1))
        self.preds, self.logits, self.B, self.t = preds, logits, B, 0
        self.seen, self.cur = [], None

    def set_online_batch(self, template, online_template):
        torch.cuda.synchronize()
        self.cur = [(GG.sha(template[b].cpu().numpy()), GG.sha(online_template[b].cpu().numpy())) for b in range(self.B)]

    def forward_test_batch(self, search, run_score_head=True):
        self.t += 1
        torch.cuda.synchronize()
        self.seen.append([self.cur[b] + (GG.sha(search[b].cpu().numpy()),) for b in range(self.B)])
        boxes = torch.tensor(np.stack([self.preds[self.t]] * self.B), device="cuda").view(-1, 1, 4)
        return {"pred_boxes": boxes, "pred_scores": torch.full((self.B,), float(self.logits[self.t]), device="cuda")}, boxes


def test_online_batched_tracker_online_size_3_matches_reference_fixture(built_lib):
    """OnlineBatchedTracker(online_size=3): growing / wrapping list of online templates, cached-template calls - states
    and the digests of every (template, online-template stack, search) handed to the network equal the run of the
    UNMODIFIED MixFormerOnline class (fixture), for both sequences of the batch."""
    from mmt_b200 import frames
    o = GG.ONLINE
    vid = GG.seeded_video(o["seed"], o["H"], o["W"], o["T"])
    preds, logits = GG.online_script(o["T"])
    logits = logits + np.float32(1.0)
    net = _ScriptedCachedNet(preds, logits, 2).cuda()
    params = types.SimpleNamespace(template_factor=o["template_factor"], template_size=o["template_size"],
                                   search_factor=o["search_factor"], search_size=o["search_size"])
    trk = frames.OnlineBatchedTracker(net, params, update_interval=2, online_size=3)
    trk.initialize([vid[0], vid[0]], [o["box"], o["box"]])
    for t in range(1, o["T"]):
        trk.track([vid[t], vid[t]])
    got = trk.results()
    for slot in (0, 1):
        assert np.array_equal(got[:, slot], GOLD["online3_states"])
        assert np.array_equal(np.array([s[slot] for s in net.seen]), GOLD["online3_inputs_sha"])
    with pytest.raises(NotImplementedError):
        trk.reset_slot(0, vid[0], o["box"])


def test_online_tracker_class_with_cached_templates(built_lib):
    """`get_tracker_class("mixformer_vit_online")` with params.online_sizes = 3 (the real model, cached templates):
    per-frame boxes equal the oracle loop whose network call is set_online + forward_test on the same model."""
    from mmt_b200 import synthetic, trackers
    cfg = synthetic.load_variant_config("mixformer_vit_online")
    params = types.SimpleNamespace(cfg=cfg, template_factor=2.0, template_size=cfg.DATA.TEMPLATE.SIZE, search_factor=4.5,
                                   search_size=cfg.DATA.SEARCH.SIZE, checkpoint=None, save_all_boxes=False,
                                   update_interval=2, online_sizes=3)
    torch.manual_seed(0)
    trk = trackers.get_tracker_class("mixformer_vit_online")(params, "lasot")
    synthetic.sharpen_(trk.network, torch.Generator().manual_seed(1000))
    trk.network.load_state_dict(trk.network.state_dict())
    T = 8
    vid = GG.seeded_video(71, 180, 240, T)
    box = [80.0, 60.0, 50.0, 40.0]
    trk.initialize(vid[0], {"init_bbox": box})
    got = [trk.track(vid[t])["target_bbox"] for t in range(1, T)]

    def net(template, online_template, search):
        trk.network.set_online(torch.from_numpy(template[None]).cuda(), torch.from_numpy(online_template).cuda())
        out, coords = trk.network.forward_test(torch.from_numpy(search[None]).cuda(), run_score_head=True)
        return coords.view(-1, 4).cpu().numpy()[0], out["pred_scores"].reshape(-1).cpu().numpy()[0]

    orc = FO.OnlineTrackerOracle(net, 2.0, params.template_size, 4.5, params.search_size, 2, online_size=3)
    orc.initialize(vid[0], box)
    for t in range(1, T):
        assert got[t - 1] == [float(v) for v in orc.track(vid[t])], t
