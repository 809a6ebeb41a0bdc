"""TEST INFRASTRUCTURE - generates tests/golden/frames_*.npz by running the UNMODIFIED reference's per-frame host glue
(from /root/reference, which only exists in the build container) on seeded synthetic frames:

  * `sample_target`                 lib/train/data/processing_utils.py:15-83
  * `Preprocessor_Multimodal`       lib/test/tracker/tracker_utils.py:37-48      (torch.Tensor.cuda shimmed to identity)
  * `MixFormer.track` box update    lib/test/tracker/asymmetric_shared_ce.py:99-103,134-140 (called unbound on a stub
    object with a fake network that returns a seeded box), `clip_box` lib/utils/box_ops.py:155-164

and asserts that oracle/frame_oracle.py reproduces every output (crops bit for bit, states exactly) before the
fixtures are written.  Run:  python oracle/gen_golden_frames.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frame_oracle as FO  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def synthetic_frame(rng, H, W):
    """Smooth-ish structured uint8 RGB frame (gradients + blobs + noise) so that interpolation matters."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.empty((H, W, 3), dtype=np.float32)
    for c in range(3):
        fx, fy, ph = rng.uniform(0.01, 0.2), rng.uniform(0.01, 0.2), rng.uniform(0, 6.28)
        img[..., c] = 127 + 90 * np.sin(xx * fx + yy * fy + ph) + rng.normal(0, 25, (H, W))
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [
    # H, W, box (x, y, w, h), note
    (480, 640, (300.0, 200.0, 60.0, 40.0), "interior"),
    (480, 640, (2.0, 3.0, 50.0, 70.0), "top-left overhang"),
    (480, 640, (600.0, 440.0, 40.0, 40.0), "bottom-right overhang (last row/column quirk)"),
    (256, 320, (100.0, 80.0, 128.5, 77.25), "fractional box, upscale of the template"),
    (720, 1280, (500.25, 300.75, 301.5, 250.0), "large target: search crop larger than the frame height"),
    (480, 640, (310.0, 230.0, 10.0, 10.0), "tiny target (margin-sized box)"),
    (300, 300, (6.0, 6.0, 288.0 / 4.5 * 2, 288.0 / 4.5 * 2), "crop exactly 2x the search size"),
    (300, 300, (100.0, 100.0, 64.0, 64.0), "crop == output size for factor 4.5 (identity resize)"),
]


def case_frames(ci):
    """The seeded (visible, infrared) uint8 frames of case ci; the fixture stores their SHA-256 instead of the pixels."""
    H, W, _, _ = CASES[ci]
    rng = np.random.default_rng(20261018 + ci)
    im_v = synthetic_frame(rng, H, W)
    im_i = synthetic_frame(rng, H, W)
    if ci % 2 == 0:
        im_i[...] = im_i[..., :1]                   # infrared frames are grey images stored with three equal channels
    return im_v, im_i                               # (odd cases keep unequal channels: the colour map must still hold)


def sha(a: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded_video(seed, H, W, T):
    """A drifting textured scene (every frame differs): list of T uint8 RGB frames."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H + 2 * T, W + 3 * T, 3), dtype=np.uint8)
    # smooth a little so that interpolation is not pure noise
    base = ((base.astype(np.uint16) + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4).astype(np.uint8)
    return [np.ascontiguousarray(base[2 * t:2 * t + H, 3 * t:3 * t + W]) for t in range(T)]


ONLINE = dict(seed=4242, H=200, W=260, T=13, box=(90.0, 70.0, 48.0, 36.0), update_interval=5, template_factor=2.0,
              template_size=128, search_factor=4.5, search_size=288)


def online_script(T):
    """Seeded (pred box, SPM logit) per frame: the stub network's outputs.  Logits are spaced well apart (no near ties)
    and exercise every branch: below 0.5, new maximum, lower than the running maximum, right after a commit."""
    rng = np.random.default_rng(99)
    preds = np.stack([rng.uniform(0.4, 0.6, T), rng.uniform(0.4, 0.6, T), rng.uniform(0.15, 0.3, T), rng.uniform(0.15, 0.3, T)], 1)
    logits = np.array([0.0, -1.0, 0.8, 0.3, 2.0, -0.5, 1.1, -2.0, -0.2, 0.6, 3.0, 0.1, 1.7][:T], dtype=np.float32)
    return preds.astype(np.float32), logits


def main_online(out):
    """One RGB sequence through the UNMODIFIED MixFormerOnline.initialize/track (lib/test/tracker/
    mixformer_convmae_online.py:62-128, online_size 1) with a stub network; asserts OnlineTrackerOracle agrees."""
    from lib.test.tracker import mixformer_convmae_online as trk
    from lib.test.tracker.tracker_utils import Preprocessor_wo_mask
    o = ONLINE
    vid = seeded_video(o["seed"], o["H"], o["W"], o["T"])
    preds, logits = online_script(o["T"])
    stub = types.SimpleNamespace()
    stub.params = types.SimpleNamespace(search_factor=o["search_factor"], search_size=o["search_size"],
                                        template_factor=o["template_factor"], template_size=o["template_size"], vis_attn=0)
    stub.preprocessor = Preprocessor_wo_mask()
    stub.online_size, stub.update_interval, stub.max_score_decay = 1, o["update_interval"], 1.0
    stub.save_all_boxes, stub.debug = False, False
    stub.cfg = None
    seen = []

    def net(template, online_template, search, run_score_head=True):
        t = stub.frame_id
        seen.append((sha(template[0].numpy()), sha(online_template[0].numpy()), sha(search[0].numpy())))
        return {"pred_boxes": torch.from_numpy(preds[t]).view(1, 1, 4), "pred_scores": torch.tensor([logits[t]])}, None

    stub.network = net
    stub.map_box_back = types.MethodType(trk.MixFormerOnline.map_box_back, stub)
    trk.MixFormerOnline.initialize(stub, vid[0], {"init_bbox": list(o["box"])})
    states, maxs = [list(o["box"])], [-1.0]
    for t in range(1, o["T"]):
        res = trk.MixFormerOnline.track(stub, vid[t])
        states.append([float(v) for v in res["target_bbox"]])
        maxs.append(float(stub.max_pred_score))
    # the oracle must reproduce states, running maxima and WHICH crops were fed to the network at every frame
    seen_o = []

    def net_o(template, online_template, search):
        t = orc.frame_id
        seen_o.append((sha(template), sha(online_template), sha(search)))
        return preds[t], logits[t]

    orc = FO.OnlineTrackerOracle(net_o, o["template_factor"], o["template_size"], o["search_factor"], o["search_size"],
                                 o["update_interval"])
    orc.initialize(vid[0], o["box"])
    for t in range(1, o["T"]):
        st = orc.track(vid[t])
        assert [float(v) for v in st] == states[t], (t, st, states[t])
        assert float(orc.max_pred_score) == maxs[t], (t, orc.max_pred_score, maxs[t])
    assert seen_o == seen
    assert len({s[1] for s in seen}) >= 3, "the online template must actually change during the sequence"
    out["online_states"] = np.array(states, dtype=np.float64)
    out["online_max_scores"] = np.array(maxs, dtype=np.float64)
    out["online_inputs_sha"] = np.array(seen)
    out["online_video_sha"] = sha(np.stack(vid))
    print("online tracker case: ok,", len({s[1] for s in seen}), "distinct online templates")


def main_online3(out):
    """online_size = 3 (the shipped ONLINE_SIZES): UNMODIFIED MixFormerOnline with a stub network whose set_online /
    forward_test record what they are given; OnlineTrackerOracle(online_size=3) must see the same template stacks."""
    from lib.test.tracker import mixformer_convmae_online as trk
    from lib.test.tracker.tracker_utils import Preprocessor_wo_mask
    o = ONLINE
    T = o["T"]
    vid = seeded_video(o["seed"], o["H"], o["W"], T)
    preds, logits = online_script(T)
    logits = logits + np.float32(1.0)                    # more frames above 0.5: the list fills up and wraps
    stub = types.SimpleNamespace()
    stub.params = types.SimpleNamespace(search_factor=o["search_factor"], search_size=o["search_size"],
                                        template_factor=o["template_factor"], template_size=o["template_size"], vis_attn=0)
    stub.preprocessor = Preprocessor_wo_mask()
    stub.online_size, stub.update_interval, stub.max_score_decay = 3, 2, 1.0
    stub.save_all_boxes, stub.debug, stub.cfg = False, False, None
    seen, cur = [], {}

    class Net:
        def set_online(self, template, online_template):
            cur["t"], cur["ot"] = sha(template[0].numpy()), sha(online_template.numpy())

        def forward_test(self, search, run_score_head=True):
            t = stub.frame_id
            seen.append((cur["t"], cur["ot"], sha(search[0].numpy())))
            return {"pred_boxes": torch.from_numpy(preds[t]).view(1, 1, 4), "pred_scores": torch.tensor([logits[t]])}, None

    stub.network = Net()
    stub.map_box_back = types.MethodType(trk.MixFormerOnline.map_box_back, stub)
    trk.MixFormerOnline.initialize(stub, vid[0], {"init_bbox": list(o["box"])})
    states = [list(o["box"])]
    for t in range(1, T):
        states.append([float(v) for v in trk.MixFormerOnline.track(stub, vid[t])["target_bbox"]])
    seen_o = []

    def net_o(template, online_template, search):
        seen_o.append((sha(template), sha(online_template), sha(search)))
        return preds[orc.frame_id], logits[orc.frame_id]

    orc = FO.OnlineTrackerOracle(net_o, o["template_factor"], o["template_size"], o["search_factor"], o["search_size"], 2,
                                 online_size=3)
    orc.initialize(vid[0], o["box"])
    for t in range(1, T):
        assert [float(v) for v in orc.track(vid[t])] == states[t], t
    assert seen_o == seen
    out["online3_states"] = np.array(states, dtype=np.float64)
    out["online3_inputs_sha"] = np.array(seen)
    print("online_size 3 case: ok,", len({s[1] for s in seen}), "distinct online-template stacks")


FULL_CASES = (0, 2)        # cases whose uint8 crops are stored in full; the others are stored as SHA-256 digests


def main():
    ref_shims.install()
    # lib/train/__init__.py and lib/train/data/__init__.py pull in the training stack (torch._six, lmdb, ...): register
    # the two packages as bare namespaces so that only processing_utils.py itself (unmodified) is executed
    for pkg in ("lib.train", "lib.train.data"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(ref_shims.REFERENCE_ROOT, *pkg.split("."))]
        sys.modules[pkg] = m
    from lib.train.data.processing_utils import sample_target
    from lib.test.tracker.tracker_utils import Preprocessor_Multimodal
    from lib.test.tracker import asymmetric_shared_ce as trk
    pre = Preprocessor_Multimodal()
    out = {}
    template_factor, template_size, search_factor, search_size = 2.0, 128, 4.5, 288
    for ci, (H, W, box, note) in enumerate(CASES):
        im_v, im_i = case_frames(ci)
        rng = np.random.default_rng(777 + ci)
        state = list(box)
        rec = {"im_v_sha": sha(im_v), "im_i_sha": sha(im_i), "box": np.array(box, dtype=np.float64)}
        for name, factor, size in (("template", template_factor, template_size), ("search", search_factor, search_size)):
            zv, rf, _ = sample_target(im_v, state, factor, output_sz=size)
            zi, rf_i, _ = sample_target(im_i, state, factor, output_sz=size)
            tv, ti = pre.process(zv, zi)
            ov, orf = FO.sample_target(im_v, state, factor, size)
            oi, _ = FO.sample_target(im_i, state, factor, size)
            assert np.array_equal(ov, zv) and np.array_equal(oi, zi), (note, name, "crop mismatch")
            assert orf == rf
            nv, ni = FO.process_multimodal(ov, oi)
            assert np.array_equal(nv, tv[0].numpy()) and np.array_equal(ni, ti[0].numpy()), (note, name, "normalise")
            # the reference's outputs: uint8 crops (after JET for the infrared one) and the normalised fp32 tensors
            zi_jet = np.rint((ti[0].numpy().transpose(1, 2, 0) * FO.STD + FO.MEAN) * 255.0).astype(np.uint8)
            assert np.array_equal(FO.normalize(zi_jet), ti[0].numpy())
            rec[f"{name}_u8_v_sha"], rec[f"{name}_u8_i_sha"] = sha(zv), sha(zi_jet)
            rec[f"{name}_v_sha"], rec[f"{name}_i_sha"] = sha(tv[0].numpy()), sha(ti[0].numpy())
            if ci in FULL_CASES:
                rec[f"{name}_u8_v"], rec[f"{name}_u8_i"] = zv, zi_jet
            rec[f"{name}_rf"] = np.float64(rf)
        # one track() step of the reference tracker class with a stub network returning seeded boxes
        preds = rng.uniform(0.05, 0.95, size=(6, 4)).astype(np.float32)
        preds[:, 2:] = rng.uniform(0.02, 0.9, size=(6, 2)).astype(np.float32)
        preds[0] = (0.5, 0.5, 0.2, 0.2)
        preds[1] = (0.99, 0.99, 0.9, 0.9)           # pushes the box out of the frame: clip_box branches
        preds[2] = (0.01, 0.01, 0.01, 0.01)
        states = []
        for p in preds:
            stub = types.SimpleNamespace()
            stub.params = types.SimpleNamespace(search_factor=search_factor, search_size=search_size,
                                                template_factor=template_factor, template_size=template_size,
                                                vis_search=0)
            stub.state = list(box)
            stub.frame_id = 0
            stub.preprocessor = pre
            stub.template = stub.online_template = None
            stub.update_intervals = [10 ** 9]
            stub.debug = False
            stub.save_all_boxes = False
            stub.network = lambda t, ot, s, return_features=False, _p=p: ({"pred_boxes": torch.from_numpy(_p).view(1, 1, 4)}, None)
            stub.map_box_back = types.MethodType(trk.MixFormer.map_box_back, stub)
            res = trk.MixFormer.track(stub, [im_v, im_i])
            got = FO.update_state(list(box), p, rec["search_rf"], search_size, H, W, margin=10)
            assert [float(v) for v in res["target_bbox"]] == [float(v) for v in got], (note, p, res, got)
            states.append([float(v) for v in res["target_bbox"]])
        rec["pred_boxes"] = preds
        rec["next_states"] = np.array(states, dtype=np.float64)
        for k, v in rec.items():
            out[f"c{ci}_{k}"] = v
        print(f"case {ci} ({note}): ok, search rf {rec['search_rf']:.6f}")
    main_online(out)
    main_online3(out)
    out["n_cases"] = np.int64(len(CASES))
    out["params"] = np.array([template_factor, template_size, search_factor, search_size], dtype=np.float64)
    path = os.path.join(GOLDEN, "frames_rgbt.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
