"""CPU: the frame-side oracle (oracle/frame_oracle.py) against
  * the fixture written by oracle/gen_golden_frames.py from the UNMODIFIED reference functions (sample_target,
    Preprocessor_Multimodal, MixFormer.track + clip_box) - tests/golden/frames_rgbt.npz,
  * OpenCV itself (the reference's third-party resize / colour-map implementation) where cv2 is importable,
and the host logic of mmt_b200/frames.py that needs no GPU."""
import os

import numpy as np
import pytest

from oracle import frame_oracle as FO
from oracle import gen_golden_frames as GG

GOLD = np.load(GG.os.path.join(GG.GOLDEN, "frames_rgbt.npz"))
N_CASES = int(GOLD["n_cases"])
T_FACTOR, T_SIZE, S_FACTOR, S_SIZE = [float(v) for v in GOLD["params"]]


def test_fixture_frames_regenerate():
    for ci in range(N_CASES):
        im_v, im_i = GG.case_frames(ci)
        assert GG.sha(im_v) == str(GOLD[f"c{ci}_im_v_sha"]) and GG.sha(im_i) == str(GOLD[f"c{ci}_im_i_sha"]), \
            "seeded frames differ from the ones the fixture was generated with: rerun oracle/gen_golden_frames.py"


@pytest.mark.parametrize("ci", range(N_CASES))
def test_oracle_crops_match_reference_fixture(ci):
    im_v, im_i = GG.case_frames(ci)
    box = GOLD[f"c{ci}_box"].tolist()
    for name, factor, size in (("template", T_FACTOR, int(T_SIZE)), ("search", S_FACTOR, int(S_SIZE))):
        cv_, rf = FO.sample_target(im_v, box, factor, size)
        ci_, _ = FO.sample_target(im_i, box, factor, size)
        ci_jet = FO.apply_jet(ci_)
        assert rf == float(GOLD[f"c{ci}_{name}_rf"])
        assert GG.sha(cv_) == str(GOLD[f"c{ci}_{name}_u8_v_sha"])          # bit-exact byte work
        assert GG.sha(ci_jet) == str(GOLD[f"c{ci}_{name}_u8_i_sha"])
        nv, ni = FO.process_multimodal(cv_, ci_)
        assert GG.sha(nv) == str(GOLD[f"c{ci}_{name}_v_sha"]) and GG.sha(ni) == str(GOLD[f"c{ci}_{name}_i_sha"])
        if ci in GG.FULL_CASES:
            assert np.array_equal(cv_, GOLD[f"c{ci}_{name}_u8_v"]) and np.array_equal(ci_jet, GOLD[f"c{ci}_{name}_u8_i"])


@pytest.mark.parametrize("ci", range(N_CASES))
def test_oracle_state_update_matches_reference_fixture(ci):
    H, W = GG.CASES[ci][0], GG.CASES[ci][1]
    box = GOLD[f"c{ci}_box"].tolist()
    rf = float(GOLD[f"c{ci}_search_rf"])
    for p, want in zip(GOLD[f"c{ci}_pred_boxes"], GOLD[f"c{ci}_next_states"]):
        got = FO.update_state(box, p, rf, int(S_SIZE), H, W, margin=10)
        assert [float(v) for v in got] == want.tolist()                    # float64 arithmetic restated exactly


def test_crop_geometry_edge_cases():
    # round-half-even of the window origin, the dropped last column, overhang on every side
    crop_sz, x1, y1, xa, xb, ya, yb = FO.crop_geometry([10.0, 10.0, 5.0, 5.0], 1.0, 100, 100)
    assert (crop_sz, x1, y1) == (5, 10, 10) and (xa, xb, ya, yb) == (10, 15, 10, 15)
    crop_sz, x1, *_ = FO.crop_geometry([10.5, 10.0, 4.0, 4.0], 1.0, 100, 100)       # 12.5 - 2 = 10.5 -> 10 (even)
    assert x1 == 10
    crop_sz, x1, *_ = FO.crop_geometry([11.5, 10.0, 4.0, 4.0], 1.0, 100, 100)       # 11.5 -> 12 (even)
    assert x1 == 12
    crop_sz, x1, y1, xa, xb, ya, yb = FO.crop_geometry([90.0, 90.0, 10.0, 10.0], 1.0, 100, 100)
    assert (x1 + crop_sz, xb, yb) == (100, 99, 99)                                    # window reaches W: column W-1 dropped
    crop_sz, x1, y1, xa, xb, ya, yb = FO.crop_geometry([5.0, 5.0, 30.0, 30.0], 2.0, 50, 40)
    assert (xa, ya) == (0, 0) and xb == 39 and yb == 49 and crop_sz == 60
    with pytest.raises(Exception):
        FO.crop_geometry([5.0, 5.0, 0.0, 3.0], 2.0, 50, 40)


def test_resize_and_colormap_match_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(60):
        sh, sw = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        dh, dw = [(128, 128), (288, 288), (192, 192), (int(rng.integers(1, 200)), int(rng.integers(1, 200)))][t % 4]
        if t % 7 == 0:
            sh, sw = 2 * dh, 2 * dw              # cv::resize switches to its 2x2 area path here; the values coincide
        if t % 11 == 0:
            sh, sw = dh, dw
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(src, (dw, dh)), FO.resize_linear_u8(src, dh, dw)), (sh, sw, dh, dw)
    g = rng.integers(0, 256, (50, 60, 3), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(g, cv2.COLOR_BGR2GRAY), FO.bgr2gray_u8(g))
    assert np.array_equal(cv2.applyColorMap(g, cv2.COLORMAP_JET), FO.apply_jet(g))
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]
    assert np.array_equal(lut, FO.jet_lut())


def test_product_lut_equals_oracle_lut(built_lib):
    from mmt_b200 import frames
    assert frames.JET_LUT_HEX == FO._JET_HEX
    assert np.array_equal(frames.jet_lut_tensor("cpu").numpy().reshape(256, 3), FO.jet_lut())


def test_batched_tracker_refuses_cpu_models(built_lib):
    import torch
    from mmt_b200 import frames
    net = torch.nn.Linear(2, 2)
    with pytest.raises(NotImplementedError):
        frames.BatchedTracker(net, params=None)


def test_result_files_have_the_reference_format(built_lib, tmp_path):
    """running.py:31-37: boxes `np.array(data).astype(int)` saved with delimiter tab / "%d"; times with "%f"."""
    from mmt_b200 import evaluation
    seq = evaluation.SequenceSpec("car1", "lasher", ["a.jpg", "b.jpg"], [1, 2, 3, 4])
    boxes = np.array([[10.9, 20.2, 30.5, 40.0], [-0.7, 5.99, 7.0, 8.49]])
    evaluation.save_tracker_output(str(tmp_path), seq, boxes, [0.25, 0.5])
    text = (tmp_path / "lasher" / "car1.txt").read_text()
    assert text == "10\t20\t30\t40\n0\t5\t7\t8\n"                       # astype(int) truncates toward zero
    assert (tmp_path / "lasher" / "car1_time.txt").read_text() == "0.250000\n0.500000\n"


def test_sequence_spec_from_reference_and_empty_shard(built_lib):
    import types
    from mmt_b200 import evaluation
    ref_seq = types.SimpleNamespace(name="s", dataset="d", frames=[("v0", "i0"), ("v1", "i1")],
                                    init_info=lambda: {"init_bbox": ([1.0, 2.0, 3.0, 4.0], [1.5, 2.5, 3.5, 4.5])})
    spec = evaluation.SequenceSpec.from_reference(ref_seq)
    assert spec.init_bbox == [1.0, 2.0, 3.0, 4.0] and spec.frames == ref_seq.frames        # the RGB box, as the tracker uses
    # rank 3 of 4 owns nothing of a 2-sequence dataset: returns without touching the (absent) GPU
    assert evaluation.run_sequences(None, None, [spec, spec], rank=3, world_size=4) == {}


def test_online_tracker_oracle_matches_reference_fixture():
    """OnlineTrackerOracle against the UNMODIFIED MixFormerOnline.initialize/track run (fixture): states, running maximum
    scores and WHICH crops reach the network at every frame (template, online template, search digests)."""
    o = GG.ONLINE
    vid = GG.seeded_video(o["seed"], o["H"], o["W"], o["T"])
    assert GG.sha(np.stack(vid)) == str(GOLD["online_video_sha"])
    preds, logits = GG.online_script(o["T"])
    seen = []

    def net(template, online_template, search):
        t = orc.frame_id
        seen.append((GG.sha(template), GG.sha(online_template), GG.sha(search)))
        return preds[t], logits[t]

    orc = FO.OnlineTrackerOracle(net, o["template_factor"], o["template_size"], o["search_factor"], o["search_size"],
                                 o["update_interval"])
    orc.initialize(vid[0], o["box"])
    for t in range(1, o["T"]):
        st = orc.track(vid[t])
        assert [float(v) for v in st] == GOLD["online_states"][t].tolist()
        assert float(orc.max_pred_score) == float(GOLD["online_max_scores"][t])
    assert np.array_equal(np.array(seen), GOLD["online_inputs_sha"])


def test_online_score_step_branches():
    m, take = FO.online_score_step(-1.0, 0.0)          # sigmoid(0) = 0.5: not > 0.5
    assert (m, take) == (-1.0, False)
    m, take = FO.online_score_step(-1.0, 0.8)
    assert take and abs(m - 0.6899744) < 1e-6
    m2, take = FO.online_score_step(m, 0.3)            # 0.574 > 0.5 but below the running maximum
    assert (m2, take) == (m, False)
    m3, take = FO.online_score_step(m, 0.3, decay=0.5)  # the maximum decays first: 0.345 < 0.574
    assert take and abs(m3 - 0.5744425) < 1e-6


class _RecordingTracker:
    """Stand-in for BatchedTracker (CPU): boxes are a function of (sequence tag, frame index) read from the frames
    themselves, so that the slot scheduler's bookkeeping can be checked exactly."""

    def __init__(self):
        import torch
        self.torch = torch
        self.frame_id = 0
        self.calls = []

    def _box(self, frame):
        tag, t = int(frame[0][0, 0, 0]), int(frame[0][0, 0, 1])
        return [float(tag), float(t), 1.0, 1.0]

    def initialize(self, frames, init_boxes, capacity_hw=None):
        self.B = len(frames)
        self.log = self.torch.zeros((4, self.B, 4), dtype=self.torch.float64)
        for b in range(self.B):
            self.log[0, b] = self.torch.tensor(init_boxes[b], dtype=self.torch.float64)
        self.calls.append(("init", self.B, capacity_hw))

    def reset_slot(self, b, frames_b, init_box):
        self.log[self.frame_id, b] = self.torch.tensor(init_box, dtype=self.torch.float64)
        self.calls.append(("reset", b, int(frames_b[0][0, 0, 0])))

    def track(self, frames, active=None):
        self.frame_id += 1
        if self.frame_id >= self.log.shape[0]:
            self.log = self.torch.cat([self.log, self.torch.zeros_like(self.log)], 0)
        live = [True] * self.B if active is None else list(active)
        for b in range(self.B):
            if live[b]:
                assert frames[b] is not None
                self.log[self.frame_id, b] = self.torch.tensor(self._box(frames[b]), dtype=self.torch.float64)
        self.calls.append(("track", tuple(live)))


def test_run_sequences_slot_scheduler(built_lib, tmp_path):
    """Host logic of the batched runner: 5 sequences of lengths 3, 1, 4, 2, 3 through 2 slots - every sequence gets its
    own frames in order, a finished slot is refilled (or goes inactive), single-frame sequences end at initialisation,
    the result files hold one row per frame."""
    from mmt_b200 import evaluation

    def frame(tag, t, hw=(6, 8)):
        f = np.zeros(hw + (3,), dtype=np.uint8)
        f[0, 0, 0], f[0, 0, 1] = tag, t
        return [f, f]

    lengths = [3, 1, 4, 2, 3]
    seqs = [evaluation.SequenceSpec(f"s{i}", "syn", [frame(i + 1, t, (6 + i, 8)) for t in range(n)], [i, i, 5, 5])
            for i, n in enumerate(lengths)]
    trk = _RecordingTracker()
    out = evaluation.run_sequences(None, None, seqs, results_dir=str(tmp_path), batch=2, tracker_factory=lambda: trk)
    assert sorted(out) == [f"s{i}" for i in range(5)]
    for i, n in enumerate(lengths):
        rows = out[f"s{i}"]
        assert rows.shape == (n, 4)
        assert rows[0].tolist() == [i, i, 5, 5]                                  # the initial box
        for t in range(1, n):
            assert rows[t].tolist() == [i + 1, t, 1.0, 1.0], (i, t, rows)        # own frames, in order
        saved = np.loadtxt(tmp_path / "syn" / f"s{i}.txt", delimiter="\t", ndmin=2)
        assert saved.shape == (n, 4)
        assert np.loadtxt(tmp_path / "syn" / f"s{i}_time.txt", ndmin=1).shape == (n,)
    assert trk.calls[0] == ("init", 2, (10, 8))                                   # capacity = largest first frame
    assert [c for c in trk.calls if c[0] == "reset"] == [("reset", 1, 3), ("reset", 0, 4), ("reset", 0, 5)]
    assert trk.calls[-1] == ("track", (True, False))                               # the tail runs with one live slot
    # sharding: rank 1 of 2 owns s1 and s3 only
    out1 = evaluation.run_sequences(None, None, seqs, batch=2, rank=1, world_size=2, tracker_factory=_RecordingTracker)
    assert sorted(out1) == ["s1", "s3"]


def test_run_sequences_resume_and_prefetch(built_lib, tmp_path):
    """skip-if-results-exist (running.py:157-171) and the decode-ahead thread pool: a second run over the same results
    directory tracks nothing; deleting one result re-tracks exactly that sequence; prefetch on / off give the same rows
    and read every frame exactly once, through the reader."""
    from mmt_b200 import evaluation
    import threading
    reads, lock = [], threading.Lock()

    def reader(path):
        tag, t = path
        with lock:
            reads.append(path)
        f = np.zeros((6, 8, 3), dtype=np.uint8)
        f[0, 0, 0], f[0, 0, 1] = tag, t
        return f

    lengths = [3, 2, 4]
    seqs = [evaluation.SequenceSpec(f"s{i}", "syn", [[(i + 1, t), (i + 1, t)] for t in range(n)], [i, i, 5, 5])
            for i, n in enumerate(lengths)]
    run = lambda **kw: evaluation.run_sequences(None, None, seqs, batch=2, tracker_factory=_RecordingTracker,
                                                reader=reader, capacity_hw=(6, 8), **kw)
    a = run(prefetch_workers=4)
    n_reads = len(reads)
    assert n_reads == 2 * sum(lengths) and len(set(reads)) == sum(lengths)       # both modalities, every frame once
    b = run(prefetch_workers=0)
    assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)
    first = run(results_dir=str(tmp_path))
    assert sorted(first) == ["s0", "s1", "s2"]
    assert run(results_dir=str(tmp_path)) == {}                                   # everything is there: nothing to do
    os.remove(tmp_path / "syn" / "s1.txt")
    again = run(results_dir=str(tmp_path))
    assert sorted(again) == ["s1"] and np.array_equal(again["s1"], first["s1"])
    assert sorted(run(results_dir=str(tmp_path), skip_existing=False)) == ["s0", "s1", "s2"]


def test_online_tracker_oracle_online_size_3_matches_reference_fixture():
    """online_size = 3: the list of online templates grows, then wraps (mixformer_convmae_online.py:115-124); states and the
    digests of (template, stacked online templates, search) per frame from the UNMODIFIED reference class."""
    o = GG.ONLINE
    vid = GG.seeded_video(o["seed"], o["H"], o["W"], o["T"])
    preds, logits = GG.online_script(o["T"])
    logits = logits + np.float32(1.0)
    seen = []

    def net(template, online_template, search):
        seen.append((GG.sha(template), GG.sha(online_template), GG.sha(search)))
        return preds[orc.frame_id], logits[orc.frame_id]

    orc = FO.OnlineTrackerOracle(net, o["template_factor"], o["template_size"], o["search_factor"], o["search_size"], 2,
                                 online_size=3)
    orc.initialize(vid[0], o["box"])
    for t in range(1, o["T"]):
        assert [float(v) for v in orc.track(vid[t])] == GOLD["online3_states"][t].tolist()
    assert np.array_equal(np.array(seen), GOLD["online3_inputs_sha"])
    assert len({s[1] for s in seen}) >= 5
