"""Seeded synthetic weights and inputs for parity tests, smoke() and bench.py (there are no checkpoints or
datasets on the box).  Deterministic for a given torch version: everything is drawn on the CPU.

Weights: the builders' default init (trunc-normal(.02) linears, default conv init, xavier fusion - the same
distributions the reference's builders use) followed by a "sharpening" pass, because at default init the 72x72
corner logits are almost flat and the soft-argmax collapses to the crop centre for any implementation
(SURVEY.md section 7): BatchNorm running stats / affine, LayerNorm affine, the fusion sampling-offset and
attention-weight projections (zero at init) and the last head convs are given seeded non-trivial values, so the
boxes depend on every stage of the forward.
Inputs: N(0,1) crops, the distribution the reference's own profiler uses (tracking/profile_model.py:164-166).
"""
from __future__ import annotations

import os

import torch

from . import builders, config

EXPERIMENTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments")

DEFAULT_YAML = {
    "mixformer_vit": "baseline",
    "mixformer_vit_rgbt": "attention_lasher_newfusion_2layer",
    "mixformer_vit_rgbt_shared": "attention_lasher_newfusion_2layer_lnspecific",
    "mixformer_vit_rgbt_unibackbone": "attention_lasher_newfusion_2layer",
    "asymmetric_shared": "attention_lasher_newfusion_2layer",
    "asymmetric_shared_ce": "attention_lasher_newfusion_2layer",
    "asymmetric_shared_online": "attention_lasher_newfusion_2layer",
    "mixformer_vit_online": "baseline",
    "mixformer_convmae_online": "baseline",
}


def load_variant_config(variant: str, yaml_name: str | None = None, overrides: dict | None = None):
    cfg = config.load_config(variant, os.path.join(EXPERIMENTS, variant, (yaml_name or DEFAULT_YAML[variant]) + ".yaml"))
    for k, v in (overrides or {}).items():
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg


def sharpen_(model: torch.nn.Module, gen: torch.Generator) -> None:
    r = lambda *s: torch.randn(*s, generator=gen)
    u = lambda *s: torch.rand(*s, generator=gen)
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, (torch.nn.BatchNorm2d, builders.FrozenBatchNorm2d)):
                n = m.weight.numel()
                m.running_mean.copy_(0.1 * r(n))
                m.running_var.copy_(0.5 + u(n))
                m.weight.copy_(1.0 + 0.2 * r(n))
                m.bias.copy_(0.1 * r(n))
            elif isinstance(m, (torch.nn.LayerNorm, torch.nn.GroupNorm)):
                m.weight.copy_(1.0 + 0.1 * r(*m.weight.shape))
                m.bias.copy_(0.05 * r(*m.bias.shape))
            elif isinstance(m, torch.nn.Linear):
                if name.endswith("sampling_offsets"):
                    m.weight.copy_(0.02 * r(*m.weight.shape))
                elif name.endswith("attention_weights"):
                    m.weight.copy_(0.05 * r(*m.weight.shape))
                    m.bias.copy_(0.1 * r(*m.bias.shape))
                else:
                    m.bias.copy_(0.02 * r(*m.bias.shape))
        head = getattr(model, "box_head", None)
        if head is not None:
            for c in ("tl", "br"):
                conv5 = getattr(head, f"conv5_{c}")
                conv5.weight.mul_(12.0)       # peaky corner distributions: soft-argmax moves off-centre
                getattr(head, f"conv4_{c}")[0].weight.mul_(2.0)


def make_model(variant: str, seed: int = 0, sharpen: bool = True, yaml_name: str | None = None,
               overrides: dict | None = None):
    """(model on CPU in eval mode, cfg).  Call .cuda() to run it."""
    cfg = load_variant_config(variant, yaml_name, overrides)
    torch.manual_seed(seed)
    model = builders.BUILDERS[variant](cfg, train=False)
    if sharpen:
        sharpen_(model, torch.Generator().manual_seed(seed + 1000))
    return model, cfg


def make_inputs(variant: str, cfg, batch: int, seed: int = 1, device="cpu", pin: bool = False):
    """(template, online_template, search): tensors for the RGB-only model, [v, i] lists for RGB-T."""
    g = torch.Generator().manual_seed(seed)
    ts, ss = cfg.DATA.TEMPLATE.SIZE, cfg.DATA.SEARCH.SIZE

    def one(size):
        t = torch.randn(batch, 3, size, size, generator=g)
        if pin:
            t = t.pin_memory()
        return t.to(device) if device != "cpu" else t

    if variant in ("mixformer_vit", "mixformer_vit_online", "mixformer_convmae_online"):
        return one(ts), one(ts), one(ss)
    t, ot, s = [one(ts), one(ts)], [one(ts), one(ts)], [one(ss), one(ss)]
    return t, ot, s


def make_online_inputs(cfg, n_online: int = 3, seed: int = 11, device="cpu"):
    """(template [1,3,T,T], online_template [n,3,T,T], search [1,3,S,S]) for the cached-template path
    (set_online / forward_test), seeded N(0,1) crops."""
    g = torch.Generator().manual_seed(seed)
    ts, ss = cfg.DATA.TEMPLATE.SIZE, cfg.DATA.SEARCH.SIZE
    t = torch.randn(1, 3, ts, ts, generator=g)
    ot = torch.randn(n_online, 3, ts, ts, generator=g)
    s = torch.randn(1, 3, ss, ss, generator=g)
    return tuple(x.to(device) if device != "cpu" else x for x in (t, ot, s))
