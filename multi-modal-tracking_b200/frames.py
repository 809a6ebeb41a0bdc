"""Batched tracker loop with the frame-side work on the device.

The reference tracker class (lib/test/tracker/asymmetric_shared_ce.py:14-140, lib/test/tracker/mixformer_vit.py) handles
ONE sequence: per frame it crops with cv2 on the host (`sample_target`), colour-maps / normalises / uploads the crop
(`Preprocessor_Multimodal`), runs the network, pulls the box back with `.tolist()` (a device synchronisation), maps it
to frame coordinates and clips it in Python.  `BatchedTracker` keeps the same `initialize` / `track` protocol and the
same arithmetic for B sequences in lock-step, but the host only uploads the raw uint8 frames: crop, resize, colour map,
normalisation, box map-back and clipping are two CUDA kernels (csrc/frames.cu) around the forward, the tracker state
lives in HBM as float64, and nothing synchronises until `results()` is read.

torch is used for device memory, streams and events only; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

# cv2.applyColorMap(np.arange(256, dtype=np.uint8)[None], cv2.COLORMAP_JET)[0] of OpenCV 4.13.0 as hex (B, G, R per
# entry): the table Preprocessor_Multimodal applies to the infrared crop (lib/test/tracker/tracker_utils.py:43).
JET_LUT_HEX = (
    "8000008400008800008c00009000009400009800009c0000a00000a40000a80000ac0000b00000b40000b80000bc0000"
    "c00000c40000c80000cc0000d00000d40000d80000dc0000e00000e40000e80000ec0000f00000f40000f80000fc0000"
    "ff0000ff0400ff0800ff0c00ff1000ff1400ff1800ff1c00ff2000ff2400ff2800ff2c00ff3000ff3400ff3800ff3c00"
    "ff4000ff4400ff4800ff4c00ff5000ff5400ff5800ff5c00ff6000ff6400ff6800ff6c00ff7000ff7400ff7800ff7c00"
    "ff8000ff8400ff8800ff8c00ff9000ff9400ff9800ff9c00ffa000ffa400ffa800ffac00ffb000ffb400ffb800ffbc00"
    "ffc000ffc400ffc800ffcc00ffd000ffd400ffd800ffdc00ffe000ffe400ffe800ffec00fff000fff400fff800fffc00"
    "feff02faff06f6ff0af2ff0eeeff12eaff16e6ff1ae2ff1edeff22daff26d6ff2ad2ff2eceff32caff36c6ff3ac2ff3e"
    "beff42baff46b6ff4ab2ff4eaeff52aaff56a6ff5aa2ff5e9eff629aff6696ff6a92ff6e8eff728aff7686ff7a82ff7e"
    "7eff827aff8676ff8a72ff8e6eff926aff9666ff9a62ff9e5effa25affa656ffaa52ffae4effb24affb646ffba42ffbe"
    "3effc23affc636ffca32ffce2effd22affd626ffda22ffde1effe21affe616ffea12ffee0efff20afff606fffa01fffe"
    "00fcff00f8ff00f4ff00f0ff00ecff00e8ff00e4ff00e0ff00dcff00d8ff00d4ff00d0ff00ccff00c8ff00c4ff00c0ff"
    "00bcff00b8ff00b4ff00b0ff00acff00a8ff00a4ff00a0ff009cff0098ff0094ff0090ff008cff0088ff0084ff0080ff"
    "007cff0078ff0074ff0070ff006cff0068ff0064ff0060ff005cff0058ff0054ff0050ff004cff0048ff0044ff0040ff"
    "003cff0038ff0034ff0030ff002cff0028ff0024ff0020ff001cff0018ff0014ff0010ff000cff0008ff0004ff0000ff"
    "0000fc0000f80000f40000f00000ec0000e80000e40000e00000dc0000d80000d40000d00000cc0000c80000c40000c0"
    "0000bc0000b80000b40000b00000ac0000a80000a40000a000009c00009800009400009000008c000088000084000080")


def jet_lut_tensor(device) -> torch.Tensor:
    return torch.frombuffer(bytearray(bytes.fromhex(JET_LUT_HEX)), dtype=torch.uint8).to(device)


class FrameUploader:
    """Raw uint8 frames of one step -> one pinned staging buffer -> one H2D copy on a side stream.

    Every image slot (modality-major, image m*B + b) owns a fixed-capacity region, so the device pointer table is built
    once; the (H, W, pitch) table changes only when a slot switches to a sequence of another frame size.  Two
    staging/device buffer pairs alternate so that the upload of step t+1 overlaps the forward of step t."""

    def __init__(self, shapes, device, capacity_hw=None):
        # shapes: list over images of (H, W); capacity_hw: (H, W) every slot must be able to hold (default: its own)
        self.device = device
        self.shapes = [tuple(int(v) for v in sh) for sh in shapes]
        cap = [max(H * W, (capacity_hw[0] * capacity_hw[1]) if capacity_hw else 0) * 3 for (H, W) in self.shapes]
        self.capacity = [(c + 255) // 256 * 256 for c in cap]          # 256-byte aligned regions
        self.offsets = [int(v) for v in np.concatenate([[0], np.cumsum(self.capacity)[:-1]])]
        self.total = int(sum(self.capacity))
        self.pinned = [torch.empty(self.total, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.dev = [torch.empty(self.total, dtype=torch.uint8, device=device) for _ in range(2)]
        self.ptrs = [torch.tensor([d.data_ptr() + o for o in self.offsets], dtype=torch.int64, device=device)
                     for d in self.dev]
        self._dims_host = torch.tensor([[H, W, W * 3] for (H, W) in self.shapes], dtype=torch.int32).pin_memory()
        self.dims = self._dims_host.to(device)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self._used = [False, False]
        self.k = 0
        self.h2d_bytes = 0
        import os
        from concurrent.futures import ThreadPoolExecutor
        n_thr = min(8, os.cpu_count() or 1)
        self._pool = ThreadPoolExecutor(max_workers=n_thr) if n_thr > 1 else None

    def set_shape(self, idx: int, H: int, W: int) -> None:
        """Image slot idx now carries H x W frames (a new sequence entered the slot).  Stream-ordered on the current
        stream: kernels already enqueued keep seeing the old table."""
        if H * W * 3 > self.capacity[idx]:
            raise ValueError(f"frame {H}x{W} exceeds the slot capacity ({self.capacity[idx]} bytes): construct the "
                             "uploader with a larger capacity_hw")
        self.shapes[idx] = (int(H), int(W))
        # a fresh pinned row per update: the asynchronous copy below may still be reading the previous one
        row = torch.tensor([[H, W, W * 3]], dtype=torch.int32).pin_memory()
        self._dims_host[idx] = row[0]
        self.dims[idx:idx + 1].copy_(row, non_blocking=True)
        self._keep = getattr(self, "_keep", [])[-64:] + [row]

    def upload(self, images, skip=None) -> int:
        """images: list over image slots of uint8 HWC numpy arrays (entries of slots listed in `skip` are ignored and
        may be None).  Returns the buffer index; the CURRENT stream waits for the copy."""
        k = self.k
        self.k ^= 1
        if self._used[k]:
            self.copied[k].synchronize()                      # host may overwrite the staging buffer
        host = self.pinned[k].numpy()
        nbytes = 0
        jobs, spans = [], []
        for i, (im, (H, W), o) in enumerate(zip(images, self.shapes, self.offsets)):
            if skip is not None and skip[i]:
                continue
            if im.shape != (H, W, 3) or im.dtype != np.uint8:
                raise ValueError(f"frame of shape {im.shape}/{im.dtype}, expected uint8 {(H, W, 3)} for image slot {i}")
            jobs.append((host[o:o + H * W * 3].reshape(H, W, 3), im))
            spans.append((o, H * W * 3))
            nbytes += H * W * 3
        # the packing memcpy (~1 MB per frame) is the host-side cost of a step: numpy releases the GIL while copying,
        # so a few threads bring it well under the device step time
        if len(jobs) >= 8 and self._pool is not None:
            list(self._pool.map(lambda j: np.copyto(j[0], j[1]), jobs))
        else:
            for dst, src in jobs:
                np.copyto(dst, src)
        with torch.cuda.stream(self.copy_stream):
            if self._used[k]:
                self.copy_stream.wait_event(self.consumed[k])  # kernels of two steps ago have read the device buffer
            if skip is None:
                self.dev[k].copy_(self.pinned[k], non_blocking=True)          # every slot is live: one copy
            else:
                for o, n in spans:                                             # only the live slots' regions travel
                    self.dev[k][o:o + n].copy_(self.pinned[k][o:o + n], non_blocking=True)
            self.copied[k].record(self.copy_stream)
        torch.cuda.current_stream().wait_event(self.copied[k])
        self._used[k] = True
        self.h2d_bytes = nbytes
        return k

    def release(self, k: int) -> None:
        """Call after the last kernel reading device buffer k has been enqueued on the current stream."""
        self.consumed[k].record(torch.cuda.current_stream())


class BatchedTracker:
    """`initialize` / `track` of the reference tracker class for B sequences at a time.

    network: a model from mmt_b200.builders on a CUDA device (RGB-T: called with [v, i] lists; RGB-only: tensors).
    params:  object with template_factor, template_size, search_factor, search_size (lib/test/parameter/*.py).
    update_intervals: the reference's `self.update_intervals` (online template refreshed when frame_id % interval == 0).
    n_mod: 2 for the RGB-T trackers, 1 for RGB-only.  jet_mask: which modalities get the JET colour map (see below).
    use_template_cache: run `cache_templates()` at template updates and `forward_search()` per frame (symmetric
    variants; bit-identical boxes).
    """

    MARGIN = 10.0      # clip_box(..., margin=10), asymmetric_shared_ce.py:103

    def __init__(self, network, params, update_intervals=(), n_mod=2, use_template_cache=False, capacity=1024,
                 jet_mask=None):
        self.network = network
        self.device = next(network.parameters()).device
        if self.device.type != "cuda":
            raise NotImplementedError("BatchedTracker needs a CUDA model (no CPU fallback)")
        self.params = params
        self.update_intervals = [int(u) for u in update_intervals]
        self.n_mod = int(n_mod)
        # bit m set: modality m goes through the JET colour map.  Default = Preprocessor_Multimodal (infrared only), the
        # preprocessor of the shared / unibackbone / asymmetric trackers; the two-stream tracker
        # (lib/test/tracker/mixformer_vit_rgbt.py:25,85-86) uses Preprocessor_wo_mask for both modalities: jet_mask=0.
        self.jet_mask = (0b10 if self.n_mod == 2 else 0) if jet_mask is None else int(jet_mask)
        self.use_cache = bool(use_template_cache)
        self.capacity = int(capacity)
        self.frame_id = 0
        self.B = 0
        self._lut = jet_lut_tensor(self.device)

    # ------------------------------------------------------------------ helpers
    def _flatten(self, frames):
        """list over sequences of [im_v, im_i] (or a single array) -> modality-major list of images."""
        if self.n_mod == 1:
            return [f if (f is None or isinstance(f, np.ndarray)) else f[0] for f in frames]
        return [f[m] for m in range(self.n_mod) for f in frames]

    def _model_args(self, buf):
        return [buf[m] for m in range(self.n_mod)] if self.n_mod > 1 else buf[0]

    def _crop(self, k, factor, size, out, active=None, rf=None):
        ops.frame_crop(self.up.ptrs[k], self.up.dims, self.state, factor, size, self.n_mod, self.jet_mask, self._lut,
                       active=active, out=out, resize_factor=rf)

    # ------------------------------------------------------------------ protocol
    def initialize(self, frames, init_boxes, capacity_hw=None):
        """frames: list over B sequences of [image_v, image_i] uint8 HWC RGB arrays (RGB-only: one array each);
        init_boxes: [B, 4] (x, y, w, h).  asymmetric_shared_ce.py:50-72: template = online template = crop at the
        initial box, state = initial box.  capacity_hw: largest (H, W) a slot may be switched to by reset_slot()."""
        p = self.params
        B = self.B = len(frames)
        imgs = self._flatten(frames)
        self.up = FrameUploader([im.shape[:2] for im in imgs], self.device, capacity_hw)
        dev = self.device
        self.state = torch.tensor(np.asarray(init_boxes, dtype=np.float64).reshape(B, 4), device=dev)
        self.rf = torch.empty(B, dtype=torch.float64, device=dev)
        self.template = torch.empty((self.n_mod, B, 3, p.template_size, p.template_size), dtype=torch.float32, device=dev)
        self.online_template = torch.empty_like(self.template)
        self.search = torch.zeros((self.n_mod, B, 3, p.search_size, p.search_size), dtype=torch.float32, device=dev)
        self.log = torch.zeros((self.capacity, B, 4), dtype=torch.float64, device=dev)
        self.frame_id = 0                       # steps taken (row of the result table)
        self.frame_ids = np.zeros(B, dtype=np.int64)     # per slot: frames tracked since its sequence was initialised
        k = self.up.upload(imgs)
        self._crop(k, float(p.template_factor), int(p.template_size), self.template)
        self.up.release(k)
        self.online_template.copy_(self.template)
        self.log[0].copy_(self.state)
        if self.use_cache:
            self.network.cache_templates(self._model_args(self.template), self._model_args(self.online_template))

    def reset_slot(self, b, frames_b, init_box):
        """Slot b starts a NEW sequence (the batched harness refills finished slots): state = init box, template =
        online template = crop of the given first frame; the other slots are untouched.  The slot's next track() frame
        is frame 1 of the new sequence; its initial box is written to the current row of the result table."""
        if self.use_cache:
            raise NotImplementedError("reset_slot with use_template_cache: the template cache is per batch")
        p = self.params
        imgs = [frames_b] if self.n_mod == 1 else list(frames_b)
        for m, im in enumerate(imgs):
            self.up.set_shape(m * self.B + b, im.shape[0], im.shape[1])
        one = torch.zeros(self.B, dtype=torch.uint8)
        one[b] = 1
        act = one.to(self.device)
        self.state[b:b + 1].copy_(torch.tensor(np.asarray(init_box, dtype=np.float64).reshape(1, 4)))
        full = [None] * (self.n_mod * self.B)
        skip = [True] * (self.n_mod * self.B)
        for m, im in enumerate(imgs):
            full[m * self.B + b], skip[m * self.B + b] = im, False
        k = self.up.upload(full, skip=skip)
        self._crop(k, float(p.template_factor), int(p.template_size), self.template, active=act)
        self.up.release(k)
        self.online_template[:, b].copy_(self.template[:, b])
        self.log[self.frame_id, b].copy_(self.state[b])
        self.frame_ids[b] = 0

    def track(self, frames, active=None):
        """One frame for every (active) sequence; returns nothing and does not synchronise (read `results()`).
        active: optional [B] bools - inactive slots keep state, crops and templates (their `frames` entry may be None)."""
        p = self.params
        self.frame_id += 1
        if self.frame_id >= self.log.shape[0]:
            self.log = torch.cat([self.log, torch.zeros_like(self.log)], 0)
        act, act_np, skip = None, None, None
        if active is not None:
            act_np = np.ascontiguousarray(np.asarray(active, dtype=np.uint8))
            act = torch.from_numpy(act_np).to(self.device)
            skip = [not act_np[i % self.B] for i in range(self.n_mod * self.B)]
        live = np.ones(self.B, dtype=bool) if act_np is None else act_np.astype(bool)
        self.frame_ids[live] += 1
        imgs = self._flatten([f if f is not None else [None] * self.n_mod for f in frames]) if active is not None \
            else self._flatten(frames)
        k = self.up.upload(imgs, skip=skip)
        self._crop(k, float(p.search_factor), int(p.search_size), self.search, active=act, rf=self.rf)
        with torch.inference_mode():
            if self.use_cache:
                _, coords = self.network.forward_search(self._model_args(self.search))
            else:
                _, coords = self.network(self._model_args(self.template), self._model_args(self.online_template),
                                         self._model_args(self.search))
        ops.track_update(coords.view(-1, 4), self.rf, self.up.dims, self.state, int(p.search_size), self.MARGIN,
                         log=self.log[self.frame_id], active=act)
        # online-template refresh: `for update_i in update_intervals: if frame_id % update_i == 0` per slot
        # (asymmetric_shared_ce.py:106-114), from the frame just tracked at the NEW state
        updated = False
        for interval in self.update_intervals:
            due = live & (self.frame_ids % interval == 0)
            if due.any():
                due_dev = None if due.all() else torch.from_numpy(due.astype(np.uint8)).to(self.device)
                self._crop(k, float(p.template_factor), int(p.template_size), self.online_template, active=due_dev)
                updated = True
        self.up.release(k)
        if updated and self.use_cache:
            self.network.cache_templates(self._model_args(self.template), self._model_args(self.online_template))

    def results(self) -> np.ndarray:
        """[frame_id + 1, B, 4] float64 boxes (x, y, w, h) per step; row 0 is the initial box.  Synchronises."""
        return self.log[:self.frame_id + 1].cpu().numpy()


class OnlineBatchedTracker(BatchedTracker):
    """`MixFormerOnline` (lib/test/tracker/mixformer_convmae_online.py:12-140, mixformer_vit_online.py) for B sequences
    with `online_size == 1`: every frame runs the full forward with the SPM score head; the crop at the new box becomes
    the online-template CANDIDATE when its score beats 0.5 and the running maximum (`mmt_online_score_update` + a masked
    `mmt_frame_crop`, no score read-back); every `update_interval` frames the candidate replaces the online template and
    the bookkeeping restarts from the first template.  `online_size > 1`: the candidate is appended to (then cycled
    through) a list of online templates and every frame runs `forward_test` against the cached templates - the
    reference does this for ONE sequence (mixformer_vit/mixformer.py:238); here `set_online_batch` /
    `forward_test_batch` carry B lock-step sequences."""

    def __init__(self, network, params, update_interval=200, max_score_decay=1.0, capacity=1024, n_mod=1, jet_mask=None,
                 online_size=1):
        # n_mod = 2: the RGB-T online tracker (lib/test/tracker/asymmetric_shared_online.py:62-119: the same bookkeeping
        # on [v, i] crop pairs, Preprocessor_Multimodal, no score decay)
        super().__init__(network, params, update_intervals=(), n_mod=n_mod, use_template_cache=False, capacity=capacity,
                         jet_mask=jet_mask)
        self.update_interval = int(update_interval)
        self.max_score_decay = float(max_score_decay)
        # online_size > 1 (mixformer_convmae_online.py:66-68,94-97,115-124): the online templates form a list that grows
        # to online_size and is then overwritten round-robin; every frame runs the search tokens only against the cached
        # templates (`set_online_batch` at every list change, `forward_test_batch` per frame).  Lock-step batches only.
        self.online_size = int(online_size)
        if self.online_size > 1 and n_mod != 1:
            raise NotImplementedError("online_size > 1 exists for the RGB-only online trackers (the reference's RGB-T "
                                      "set_online is broken, asymmetric_shared_online.py:386-387)")

    def initialize(self, frames, init_boxes, capacity_hw=None):
        super().initialize(frames, init_boxes, capacity_hw)
        self.max_score = torch.full((self.B,), -1.0, dtype=torch.float64, device=self.device)
        self.take = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
        self.online_max_template = self.template.clone()
        self.scores = torch.zeros((self.log.shape[0], self.B), dtype=torch.float32, device=self.device)
        if self.online_size > 1:
            self.online_stack = self.template[0].unsqueeze(1).clone()        # [B, 1, 3, T, T]: online_template = template
            self.online_forget_id = 0
            self.network.set_online_batch(self.template[0], self.online_stack)

    def reset_slot(self, b, frames_b, init_box):
        if self.online_size > 1:
            raise NotImplementedError("online_size > 1 runs lock-step batches (one online-template list length for all slots)")
        super().reset_slot(b, frames_b, init_box)
        self.max_score[b:b + 1].fill_(-1.0)
        self.online_max_template[:, b].copy_(self.template[:, b])

    def track(self, frames, active=None):
        p = self.params
        self.frame_id += 1
        if self.frame_id >= self.log.shape[0]:
            self.log = torch.cat([self.log, torch.zeros_like(self.log)], 0)
            self.scores = torch.cat([self.scores, torch.zeros_like(self.scores)], 0)
        act, act_np, skip = None, None, None
        if active is not None:
            act_np = np.ascontiguousarray(np.asarray(active, dtype=np.uint8))
            act = torch.from_numpy(act_np).to(self.device)
            skip = [not act_np[i % self.B] for i in range(self.n_mod * self.B)]
        live = np.ones(self.B, dtype=bool) if act_np is None else act_np.astype(bool)
        self.frame_ids[live] += 1
        imgs = self._flatten([f if f is not None else [None] * self.n_mod for f in frames]) if active is not None \
            else self._flatten(frames)
        k = self.up.upload(imgs, skip=skip)
        self._crop(k, float(p.search_factor), int(p.search_size), self.search, active=act, rf=self.rf)
        if self.online_size > 1 and active is not None:
            raise NotImplementedError("online_size > 1 runs lock-step batches")
        with torch.inference_mode():
            if self.online_size > 1:
                out, coords = self.network.forward_test_batch(self.search[0], run_score_head=True)
            else:
                out, coords = self.network(self._model_args(self.template), self._model_args(self.online_template),
                                           self._model_args(self.search), run_score_head=True)
        logits = out["pred_scores"].reshape(-1).contiguous()
        self.scores[self.frame_id].copy_(logits)
        ops.track_update(coords.view(-1, 4), self.rf, self.up.dims, self.state, int(p.search_size), self.MARGIN,
                         log=self.log[self.frame_id], active=act)
        ops.online_score_update(logits, self.max_score, self.take, self.max_score_decay, active=act)
        self._crop(k, float(p.template_factor), int(p.template_size), self.online_max_template, active=self.take)
        self.up.release(k)
        due = live & (self.frame_ids % self.update_interval == 0)
        if due.all() and self.online_size > 1:
            cand = self.online_max_template[0]
            if self.online_stack.shape[1] < self.online_size:
                self.online_stack = torch.cat([self.online_stack, cand.unsqueeze(1)], dim=1)
            else:
                self.online_stack[:, self.online_forget_id].copy_(cand)
                self.online_forget_id = (self.online_forget_id + 1) % self.online_size
            self.network.set_online_batch(self.template[0], self.online_stack)
            self.online_max_template.copy_(self.template)
            self.max_score.fill_(-1.0)
        elif due.all():
            self.online_template.copy_(self.online_max_template)
            self.online_max_template.copy_(self.template)
            self.max_score.fill_(-1.0)
        else:
            for b in np.nonzero(due)[0]:
                b = int(b)
                self.online_template[:, b].copy_(self.online_max_template[:, b])
                self.online_max_template[:, b].copy_(self.template[:, b])
                self.max_score[b:b + 1].fill_(-1.0)

    def score_logits(self) -> np.ndarray:
        """[frame_id + 1, B] raw SPM logits per step (row 0 unused).  Synchronises."""
        return self.scores[:self.frame_id + 1].cpu().numpy()
