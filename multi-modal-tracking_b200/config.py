"""Config objects for the tracker builders: per-variant defaults + YAML overlay.

Mirrors the reference's three-layer config system for the keys the forward reads
(lib/config/<script>/config.py defaults; experiments/<script>/<name>.yaml overlay through
`update_config_from_file`, unknown keys raise ValueError as in config.py:124-135; RGB-T test-time overlay
experiments/tracking.yaml, lib/test/parameter/asymmetric_shared_ce.py:12-15).  The reference's YAML files are
consumed unchanged.  TRAIN.* keys are carried verbatim but not validated: training is out of scope.
"""
from __future__ import annotations

import copy

import yaml


class Cfg(dict):
    """Attribute-style nested dict (stand-in for easydict.EasyDict, which the image does not ship)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Cfg):
            v = Cfg(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})


_UPDATE_INTERVALS = {"LASOT": [200], "GOT10K_TEST": [200], "TRACKINGNET": [200], "VOT20": [200], "VOT20LT": [200]}

_BASE = {
    "MODEL": {
        "VIT_TYPE": "base_patch16",
        "HEAD_TYPE": "CORNER",
        "HIDDEN_DIM": 768,
        "NUM_OBJECT_QUERIES": 1,
        "POSITION_EMBEDDING": "sine",
        "PREDICT_MASK": False,
        "BACKBONE": {"PRETRAINED": True, "PRETRAINED_PATH": ""},
        "FUSION_LAYERS": 6,
    },
    "TRAIN": {},
    "DATA": {
        "SAMPLER_MODE": "causal",
        "MEAN": [0.485, 0.456, 0.406],
        "STD": [0.229, 0.224, 0.225],
        "MAX_SAMPLE_INTERVAL": [200],
        "TRAIN": {"DATASETS_NAME": ["GOT10K_vottrain"], "DATASETS_RATIO": [1], "SAMPLE_PER_EPOCH": 60000},
        "VAL": {"DATASETS_NAME": ["GOT10K_votval"], "DATASETS_RATIO": [1], "SAMPLE_PER_EPOCH": 10000},
        "SEARCH": {"SIZE": 288, "FACTOR": 5.0, "CENTER_JITTER": 4.5, "SCALE_JITTER": 0.5},
        "TEMPLATE": {"SIZE": 128, "FACTOR": 2.0, "NUMBER": 1, "CENTER_JITTER": 0, "SCALE_JITTER": 0},
    },
    "TEST": {
        "TEMPLATE_FACTOR": 2.0,
        "TEMPLATE_SIZE": 128,
        "SEARCH_FACTOR": 5.0,
        "SEARCH_SIZE": 288,
        "EPOCH": 500,
        "UPDATE_INTERVALS": dict(_UPDATE_INTERVALS),
        "LOAD_FROME_TRAIN_RESULT": False,
    },
}

_RGBT = {"MODEL": {"RGBT_PRETRAINED_PATH": "", "FUSION_CLASS": "Attention_Fusion_Bimodal"}}
_ONLINE = {
    "MODEL": {"HEAD_FREEZE_BN": False, "PRETRAINED_STAGE1": False},
    "TEST": {
        "UPDATE_INTERVALS": {**_UPDATE_INTERVALS, "OTB": [200], "UAV": [200]},
        "ONLINE_SIZES": {k: [3] for k in ("LASOT", "GOT10K_TEST", "TRACKINGNET", "VOT20", "VOT20LT", "OTB", "UAV")},
    },
}

# variant -> (overlays, keys removed from the base tree)
_VARIANTS = {
    "mixformer_vit": ([{"MODEL": {"RGB_PRETRAINED_PATH": ""}, "DATA": {"SAMPLER_MODE": "casual",
                                                                         "MAX_SAMPLE_INTERVAL": 200}}], []),
    "mixformer_vit_rgbt": ([_RGBT], []),
    "mixformer_vit_rgbt_shared": ([_RGBT], []),
    "mixformer_vit_rgbt_unibackbone": ([_RGBT], []),
    "asymmetric_shared": ([_RGBT], []),
    "asymmetric_shared_ce": ([_RGBT, {"MODEL": {"BACKBONE": {"STRIDE": 16, "CE_LOC": [3, 6, 9],
                                                               "CE_KEEP_RATIO": [0.7, 0.7, 0.7],
                                                               "CE_TEMPLATE_RANGE": "CTR_POINT"}}}], []),
    # lib/config/asymmetric_shared_online/config.py: the RGB-T tree with TRACKER_/SCORE_PRETRAINED_PATH instead of
    # RGBT_PRETRAINED_PATH (no ONLINE_SIZES: the tracker class falls back to online_size 3 / its own defaults)
    "asymmetric_shared_online": ([_RGBT, {"MODEL": {"TRACKER_PRETRAINED_PATH": "", "SCORE_PRETRAINED_PATH": ""}}],
                                 [("MODEL", "RGBT_PRETRAINED_PATH")]),
    "mixformer_vit_online": ([_ONLINE], [("MODEL", "FUSION_LAYERS"), ("TEST", "LOAD_FROME_TRAIN_RESULT")]),
    "mixformer_convmae_online": ([_ONLINE, {"MODEL": {"VIT_TYPE": "convmae_base"}}],
                                 [("MODEL", "FUSION_LAYERS"), ("TEST", "LOAD_FROME_TRAIN_RESULT")]),
}


def _merge(dst: dict, src: dict) -> None:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)


def default_config(variant: str) -> Cfg:
    """Default config tree of `variant` (= the reference's lib/config/<variant>/config.py:cfg)."""
    if variant not in _VARIANTS:
        raise KeyError(f"unknown tracker variant {variant!r}; known: {sorted(_VARIANTS)}")
    tree = copy.deepcopy(_BASE)
    overlays, removed = _VARIANTS[variant]
    for o in overlays:
        _merge(tree, o)
    for sect, key in removed:
        tree[sect].pop(key, None)
    return Cfg(tree)


def _update_config(base: Cfg, exp: dict, path: str = "") -> None:
    for k, v in exp.items():
        if path == "TRAIN":           # carried, not validated (training is out of scope)
            base[k] = v
            continue
        if k not in base:
            raise ValueError("{} not exist in config.py".format(path + "." + k if path else k))
        if isinstance(v, dict) and isinstance(base[k], dict):
            _update_config(base[k], v, k if not path else path + "." + k)
        else:
            base[k] = v


def update_config_from_file(cfg: Cfg, filename: str) -> Cfg:
    with open(filename) as f:
        exp = yaml.safe_load(f) or {}
    _update_config(cfg, exp)
    return cfg


def load_config(variant: str, *yaml_files: str) -> Cfg:
    cfg = default_config(variant)
    for y in yaml_files:
        update_config_from_file(cfg, y)
    return cfg
