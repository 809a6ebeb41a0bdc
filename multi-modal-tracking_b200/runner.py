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

    step(template, online_template, search) takes CPU tensors (RGB-only) or [v, i] lists of CPU tensors (RGB-T)
    shaped like the model's forward arguments, stages them through pinned memory when they are not already
    pinned, copies them to the device, runs the model and returns the [B, 4] cxcywh boxes as a pinned CPU tensor
    after synchronising.  Nothing is cached between steps.  For RGB-T models the copies run on a side stream, one
    event per modality, so that the thermal crops travel while the RGB stream is already being embedded.
    """

    def __init__(self, model, device=None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise NotImplementedError("FrameStep needs a CUDA model (no CPU fallback)")
        self._dev = {}
        self._pin = {}
        self._out = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._events = [torch.cuda.Event(), torch.cuda.Event()]

    def _to_device(self, key, t: torch.Tensor) -> torch.Tensor:
        if t.is_cuda:
            raise ValueError("FrameStep.step takes host tensors (use the model directly for device tensors)")
        d = self._dev.get(key)
        if d is None or d.shape != t.shape:
            d = torch.empty(t.shape, dtype=torch.float32, device=self.device)
            self._dev[key] = d
        if not t.is_pinned():
            p = self._pin.get(key)
            if p is None or p.shape != t.shape:
                p = torch.empty(t.shape, dtype=torch.float32).pin_memory()
                self._pin[key] = p
            p.copy_(t)
            t = p
        d.copy_(t, non_blocking=True)
        self.h2d_bytes += t.numel() * t.element_size()
        return d

    def step(self, template, online_template, search) -> torch.Tensor:
        self.h2d_bytes = self.d2h_bytes = 0
        rgbt = isinstance(search, (list, tuple))
        with torch.cuda.device(self.device):
            if rgbt:
                cur = torch.cuda.current_stream()
                self._copy_stream.wait_stream(cur)         # the device buffers are free again (previous step is done)
                args = ([None, None], [None, None], [None, None])
                with torch.cuda.stream(self._copy_stream):
                    for m in range(2):
                        for k, (name, a) in enumerate((("t", template), ("ot", online_template), ("s", search))):
                            args[k][m] = self._to_device((name, m), a[m])
                        self._events[m].record(self._copy_stream)
                out, coords = self.model(*args, ready_events=self._events)
            else:
                args = [self._to_device((name, 0), a) for name, a in
                        (("t", template), ("ot", online_template), ("s", search))]
                out, coords = self.model(*args)
            boxes = coords.view(-1, 4)
            if self._out is None or self._out.shape != boxes.shape:
                self._out = torch.empty(boxes.shape, dtype=torch.float32).pin_memory()
            self._out.copy_(boxes, non_blocking=True)
            self.d2h_bytes = boxes.numel() * 4
            torch.cuda.current_stream().synchronize()
        return self._out
