"""Single-sequence tracker classes with the reference's interface, on the device-side frame path.

The reference ships one tracker module per variant under lib/test/tracker/ (`mixformer_vit.py`, `mixformer_vit_rgbt.py`,
`mixformer_vit_rgbt_shared.py`, `mixformer_vit_rgbt_unibackbone.py`, `asymmetric_shared.py`, `asymmetric_shared_ce.py`,
`mixformer_vit_online.py`, `mixformer_convmae_online.py`), each exporting `get_tracker_class()` whose class is built as
`cls(params, dataset_name)` and driven by `initialize(image, info)` / `track(image, info)` returning
`{"target_bbox": [x, y, w, h]}` (lib/test/evaluation/tracker_rgbt.py:100-184).  `get_tracker_class(variant)` returns a
class with exactly that protocol whose per-frame work is `frames.BatchedTracker` with B = 1: the uint8 frame is uploaded
raw, crop / resize / colour map / normalisation / forward / map-back / clip run on the device, and only the four
float64 state values come back (the API returns the box every frame, so one small D2H + synchronise per frame remains).
Same arithmetic, hence the same boxes, as the reference classes (tests/test_frames_gpu.py).

`params` carries what the reference's `lib/test/parameter/<variant>.py` puts there: cfg, template_factor, template_size,
search_factor, search_size, checkpoint (optional here: None keeps the builder's initialisation), save_all_boxes.
"""
from __future__ import annotations

import torch

from . import builders
from .frames import BatchedTracker, OnlineBatchedTracker

# which preprocessor the reference's tracker module of each variant uses (lib/test/tracker/<variant>.py):
# Preprocessor_Multimodal -> JET on the infrared crop (bit 1); Preprocessor_wo_mask -> none
_JET_MASK = {"mixformer_vit": 0, "mixformer_vit_rgbt": 0, "mixformer_vit_rgbt_shared": 0b10,
             "mixformer_vit_rgbt_unibackbone": 0b10, "asymmetric_shared": 0b10, "asymmetric_shared_ce": 0b10}
_ONLINE = ("mixformer_vit_online", "mixformer_convmae_online", "asymmetric_shared_online")


class _TrackerBase:
    variant = ""

    def __init__(self, params, dataset_name):
        self.params = params
        self.cfg = params.cfg
        network = builders.BUILDERS[self.variant](params.cfg, train=False)
        ckpt = getattr(params, "checkpoint", None)
        if ckpt:
            # same call as the reference classes (asymmetric_shared_ce.py:19-21): strict load of ckpt["net"]
            network.load_state_dict(torch.load(ckpt, map_location="cpu")["net"], strict=True)
        self.network = network.cuda()
        self.network.eval()
        self.state = None
        self.frame_id = 0
        self.save_all_boxes = getattr(params, "save_all_boxes", False)
        name = dataset_name.upper()
        ui = self.cfg.TEST.UPDATE_INTERVALS
        self.update_intervals = list(ui[name]) if name in ui else list(_as_list(self.cfg.DATA.MAX_SAMPLE_INTERVAL))

    def _result(self):
        self.state = [float(v) for v in self._trk.state[0].tolist()]         # D2H of 4 float64 + synchronise
        return {"target_bbox": self.state}


def _as_list(v):
    return v if isinstance(v, (list, tuple)) else [v]


class _FullForwardTracker(_TrackerBase):
    """lib/test/tracker/asymmetric_shared_ce.py:14-140 and its siblings (full forward every frame, periodic
    online-template refresh)."""

    def initialize(self, image, info: dict):
        rgbt = self.variant != "mixformer_vit"
        box = info["init_bbox"][0] if rgbt else info["init_bbox"]      # "simple using RGB bbox" (:68)
        self._trk = BatchedTracker(self.network, self.params, update_intervals=self.update_intervals,
                                   n_mod=2 if rgbt else 1, jet_mask=_JET_MASK[self.variant])
        self._trk.initialize([list(image) if rgbt else image], [list(box)])
        self.state = [float(v) for v in box]
        self.frame_id = 0
        if self.save_all_boxes:
            return {"all_boxes": list(info["init_bbox"]) * self.cfg.MODEL.NUM_OBJECT_QUERIES}

    def track(self, image, info: dict = None):
        self.frame_id += 1
        rgbt = self.variant != "mixformer_vit"
        self._trk.track([list(image) if rgbt else image])
        return self._result()


class _OnlineTracker(_TrackerBase):
    """lib/test/tracker/mixformer_convmae_online.py:12-145 / mixformer_vit_online.py (params.online_sizes = 1: full
    forward per frame; > 1: cached templates, list of online templates) and asymmetric_shared_online.py."""

    def __init__(self, params, dataset_name):
        super().__init__(params, dataset_name)
        self.update_interval = getattr(params, "update_interval", self.update_intervals[0])
        self.max_score_decay = getattr(params, "max_score_decay", 1.0)
        self.online_size = int(getattr(params, "online_sizes", 1))        # mixformer_convmae_online.py:45-46

    def initialize(self, image, info: dict):
        rgbt = self.variant == "asymmetric_shared_online"       # lib/test/tracker/asymmetric_shared_online.py
        box = info["init_bbox"][0] if rgbt else info["init_bbox"]
        self._trk = OnlineBatchedTracker(self.network, self.params, update_interval=self.update_interval,
                                         max_score_decay=self.max_score_decay, n_mod=2 if rgbt else 1,
                                         jet_mask=0b10 if rgbt else 0, online_size=1 if rgbt else self.online_size)
        self._trk.initialize([list(image) if rgbt else image], [list(box)])
        self.state = [float(v) for v in box]
        self.frame_id = 0

    def track(self, image, info: dict = None):
        self.frame_id += 1
        self._trk.track([list(image) if self.variant == "asymmetric_shared_online" else image])
        return self._result()


def get_tracker_class(variant: str):
    """The counterpart of `lib.test.tracker.<variant>.get_tracker_class()`."""
    if variant not in builders.BUILDERS:
        raise KeyError(f"unknown tracker variant {variant!r}; known: {sorted(builders.BUILDERS)}")
    base = _OnlineTracker if variant in _ONLINE else _FullForwardTracker
    return type("MixFormerOnline" if variant in _ONLINE else "MixFormer", (base,), {"variant": variant})
