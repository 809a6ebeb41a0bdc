"""Batched per-frame step with HOST buffers and sequence sharding (the caller-facing side of the forward).

The reference tracks one sequence per process and shards whole sequences over GPU workers with no
communication (lib/test/evaluation/running.py:134-141, 200-238); its per-frame loop copies the preprocessed
crops to the device, runs `network(...)` and reads the box back (lib/test/tracker/asymmetric_shared_ce.py:74-132:
Preprocessor H2D, `pred_boxes.mean(0) ... .tolist()` D2H).  `FrameStep` is that loop body for B sequences at a
time: pinned host crops -> device -> model forward -> boxes on the host.  `shard_sequences` is the reference's
worker assignment (sequence s -> worker s mod G), `gather_boxes` the one collective of the design (boxes and
timings to rank 0; never inside the forward).
"""
from __future__ import annotations

import torch


def shard_sequences(n_sequences: int, world_size: int, rank: int) -> list[int]:
    """Sequence ids owned by `rank`: round-robin, like the reference's worker pool (running.py:134-141 assigns
    worker i to GPU (i-1) % device_count and hands out whole sequences)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(range(rank, n_sequences, world_size))


def gather_boxes(boxes: torch.Tensor, group=None) -> torch.Tensor | None:
    """All-gather of the per-rank [B_local, 4] boxes (every rank must hold the same B_local); returns
    [world, B_local, 4] on every rank.  With world_size 1 / no process group this is a view."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return boxes.unsqueeze(0)
    world = dist.get_world_size(group)
    out = torch.empty((world * boxes.shape[0],) + tuple(boxes.shape[1:]), dtype=boxes.dtype, device=boxes.device)
    dist.all_gather_into_tensor(out, boxes.contiguous(), group=group)      # rank-major concatenation
    return out.view((world,) + tuple(boxes.shape))


def unshard_boxes(gathered: torch.Tensor, n_sequences: int) -> torch.Tensor:
    """[world, B_local, 4] (rank-major, round-robin sharding) -> [n_sequences, 4] in sequence order."""
    world, b_local = gathered.shape[0], gathered.shape[1]
    out = gathered.transpose(0, 1).reshape(world * b_local, -1)      # sequence s = local * world + rank
    return out[:n_sequences]


class FrameStep:
    """One tracked frame for B sequences: host crops in, host boxes out.

    step(template, online_template, search) takes CPU tensors (RGB-only) or [v, i] lists of CPU tensors (RGB-T),
    stages them through pinned memory when they are not already pinned, copies them to the device, runs the model and
    returns the [B, 4] cxcywh boxes as a pinned CPU tensor after synchronising.  Crops may be
      * fp32 [B, 3, S, S], already normalised (the model's forward arguments), or
      * uint8 [B, S, S, 3] as `sample_target` returns them: uploaded as bytes and normalised on the device by
        `mmt_preprocess_u8` - what the reference's `Preprocessor_*.process` does after its own uint8 upload
        (lib/test/tracker/tracker_utils.py:24-48); `jet_mask` bit m = modality m gets the JET colour map first
        (default: infrared only, Preprocessor_Multimodal; 0 for the two-stream tracker's Preprocessor_wo_mask).
    `set_templates()` keeps the (online) templates on the device between frames like the reference trackers do
    (`self.template` / `self.online_template` are CUDA tensors that change only at template updates); step(None, None,
    search) then uploads the search crops only.  For RGB-T models the copies run on a side stream, one event per
    modality, so that the thermal crops travel while the RGB stream is already being embedded.
    """

    def __init__(self, model, device=None, jet_mask=None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise NotImplementedError("FrameStep needs a CUDA model (no CPU fallback)")
        self._dev = {}
        self._pin = {}
        self._out = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.jet_mask = jet_mask
        self._lut = None
        self._resident = None
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._events = [torch.cuda.Event(), torch.cuda.Event()]

    def _to_device(self, key, t: torch.Tensor, modality=0) -> torch.Tensor:
        """Host tensor -> device fp32 [B, 3, S, S] on the CURRENT stream (copy, and the preprocessing of uint8 crops)."""
        if t.is_cuda:
            raise ValueError("FrameStep.step takes host tensors (use the model directly for device tensors)")
        if t.dtype not in (torch.float32, torch.uint8):
            raise ValueError(f"crops must be fp32 NCHW or uint8 NHWC, got {t.dtype}")
        d = self._dev.get(key)
        if d is None or d.shape != t.shape or d.dtype != t.dtype:
            d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            self._dev[key] = d
        if not t.is_pinned():
            p = self._pin.get(key)
            if p is None or p.shape != t.shape or p.dtype != t.dtype:
                p = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                self._pin[key] = p
            p.copy_(t)
            t = p
        d.copy_(t, non_blocking=True)
        self.h2d_bytes += t.numel() * t.element_size()
        if t.dtype == torch.float32:
            return d
        from . import ops
        B, S = t.shape[0], t.shape[1]
        f = self._dev.get((key, "f32"))
        if f is None or f.shape != (B, 3, S, S):
            f = torch.empty((B, 3, S, S), dtype=torch.float32, device=self.device)
            self._dev[(key, "f32")] = f
        mask = (0b10 if self.jet_mask is None else int(self.jet_mask)) >> modality & 1
        if mask and self._lut is None:
            from .frames import jet_lut_tensor
            self._lut = jet_lut_tensor(self.device)
        ops.preprocess_u8(d, f, B, jet_mask=mask, jet_lut=self._lut if mask else None)
        return f

    def _upload(self, named, rgbt):
        """named: [(name, host arg)] -> device args (lists for RGB-T); RGB-T copies go to the side stream, one event per
        modality."""
        if not rgbt:
            return [self._to_device((name, 0), a) for name, a in named]
        cur = torch.cuda.current_stream()
        self._copy_stream.wait_stream(cur)         # the device buffers are free again (previous step is done)
        args = [[None, None] for _ in named]
        with torch.cuda.stream(self._copy_stream):
            for m in range(2):
                for k, (name, a) in enumerate(named):
                    args[k][m] = self._to_device((name, m), a[m], modality=m)
                self._events[m].record(self._copy_stream)
        return args

    def set_templates(self, template, online_template) -> None:
        """Upload the template and online-template crops once; they stay on the device until the next call."""
        rgbt = isinstance(template, (list, tuple))
        with torch.cuda.device(self.device):
            t, ot = self._upload([("rt", template), ("rot", online_template)], rgbt)
            if rgbt:
                for e in self._events:
                    torch.cuda.current_stream().wait_event(e)
        self._resident = (t, ot)

    def step(self, template, online_template, search) -> torch.Tensor:
        self.h2d_bytes = self.d2h_bytes = 0
        rgbt = isinstance(search, (list, tuple))
        resident = template is None and online_template is None
        if resident and self._resident is None:
            raise RuntimeError("step(None, None, search) needs set_templates() first")
        with torch.cuda.device(self.device):
            if resident:
                (s,) = self._upload([("s", search)], rgbt)
                t, ot = self._resident
            else:
                t, ot, s = self._upload([("t", template), ("ot", online_template), ("s", search)], rgbt)
            if rgbt:
                out, coords = self.model(t, ot, s, ready_events=self._events)
            else:
                out, coords = self.model(t, ot, s)
            boxes = coords.view(-1, 4)
            if self._out is None or self._out.shape != boxes.shape:
                self._out = torch.empty(boxes.shape, dtype=torch.float32).pin_memory()
            self._out.copy_(boxes, non_blocking=True)
            self.d2h_bytes = boxes.numel() * 4
            torch.cuda.current_stream().synchronize()
        return self._out
