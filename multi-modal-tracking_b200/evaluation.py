"""Batched sequence runner that writes the reference's result files.

The reference evaluates with one worker PROCESS per sequence (lib/test/evaluation/running.py:134-238: `run_dataset` ->
`run_sequence` -> `Tracker._track_sequence`, lib/test/evaluation/tracker_rgbt.py:124-184) and saves, per sequence,
`<results_dir>/<dataset>/<name>.txt` (boxes, `astype(int)`, tab separated, "%d") and `<name>_time.txt` ("%f")
(`_save_tracker_output`, running.py:16-128).  The analysis scripts (tracking/analysis_results*.py, lib/test/utils/
load_text.py) read exactly those files.

Here one process per GPU keeps B sequence SLOTS busy: every step advances all live slots by one frame through
`frames.BatchedTracker` (frames uploaded raw, everything else on the device), a finished slot is refilled from the
rank's queue (`runner.shard_sequences`: sequence s -> rank s mod G, the reference's worker assignment), and the result
files are written in the reference's format when a sequence completes.  There is no collective: sequences are
independent.  Per-frame times: the reference stores each frame's wall time; a batched step has one wall time for all
its frames, so `<name>_time.txt` holds step time / live slots for every frame (amortised per-frame cost; the sum over a
sequence is its share of the run, which is what `fps.py`-style summaries need).
"""
from __future__ import annotations

import os
import time
from collections import deque

import numpy as np

from . import runner
from .frames import BatchedTracker


def read_image_rgb(path):
    """The reference's frame reader (tracker_rgbt.py `Video.read_image`): cv2.imread + BGR -> RGB."""
    import cv2
    im = cv2.imread(path)
    if im is None:
        raise FileNotFoundError(path)
    return cv2.cvtColor(im, cv2.COLOR_BGR2RGB)


class SequenceSpec:
    """What the runner needs of the reference's `Sequence` (lib/test/evaluation/data.py:23-60): name, dataset, the
    frame list (RGB-T: (visible, infrared) pairs; entries are paths or already-decoded uint8 HWC arrays) and the
    initial box (x, y, w, h) - `seq.init_info()["init_bbox"][0]` for the RGB-T datasets."""

    def __init__(self, name, dataset, frames, init_bbox):
        self.name, self.dataset, self.frames = name, dataset, list(frames)
        self.init_bbox = [float(v) for v in init_bbox]

    @classmethod
    def from_reference(cls, seq):
        box = seq.init_info().get("init_bbox")
        if isinstance(box, (list, tuple)) and len(box) and isinstance(box[0], (list, tuple, np.ndarray)):
            box = box[0]                                       # RGB-T: (bbox_v, bbox_i), the RGB one is used
        return cls(seq.name, seq.dataset, seq.frames, box)


def save_tracker_output(results_dir, seq, boxes, times):
    """running.py:16-128 for the single-object keys this path produces (`target_bbox`, `time`)."""
    base = os.path.join(results_dir, seq.dataset, seq.name)
    os.makedirs(os.path.dirname(base), exist_ok=True)
    np.savetxt(base + ".txt", np.array(boxes).astype(int), delimiter="\t", fmt="%d")
    np.savetxt(base + "_time.txt", np.array(times).astype(float), delimiter="\t", fmt="%f")


def _load(frame, reader, n_mod):
    if n_mod == 1:
        f = frame[0] if isinstance(frame, (list, tuple)) else frame
        return f if isinstance(f, np.ndarray) else reader(f)
    return [f if isinstance(f, np.ndarray) else reader(f) for f in frame]


def results_exist(results_dir, seq):
    """running.py:157-165 (`_results_exist`, single-object sequences): the box file of the sequence is already there."""
    return results_dir is not None and os.path.exists(os.path.join(results_dir, seq.dataset, seq.name + ".txt"))


def run_sequences(network, params, sequences, results_dir=None, batch=64, update_intervals=(), n_mod=2,
                  reader=read_image_rgb, rank=0, world_size=1, capacity_hw=None, tracker_factory=None,
                  skip_existing=True, prefetch_workers=8):
    """Track every sequence owned by `rank`; returns {name: [T, 4] float64 boxes} and writes the reference's result
    files when results_dir is given.  `sequences`: SequenceSpec list (or reference Sequence objects).
    skip_existing: resume like the reference (running.py:157-171): a sequence whose result file exists is not tracked
    again (and is absent from the returned dict).  prefetch_workers: frames of the NEXT step are decoded by a thread pool
    while the GPU runs the current one (the reference reads every frame synchronously inside its per-frame loop,
    tracker_rgbt.py:144-179; `track()` here never synchronises, so decode and forward overlap); 0 = decode in the loop."""
    specs = [s if isinstance(s, SequenceSpec) else SequenceSpec.from_reference(s) for s in sequences]
    mine = [specs[i] for i in runner.shard_sequences(len(specs), world_size, rank)]
    if skip_existing:
        mine = [s for s in mine if not results_exist(results_dir, s)]
    out = {}
    if not mine:
        return out
    queue = deque(mine)
    B = min(batch, len(mine))
    if capacity_hw is None:           # largest first frame of the rank's sequences (frame sizes are fixed per sequence)
        hw = [np.shape(_load(s.frames[0], reader, n_mod)[0] if n_mod > 1 else _load(s.frames[0], reader, n_mod))[:2]
              for s in mine]
        capacity_hw = (max(h for h, _ in hw), max(w for _, w in hw))
    # tracker_factory: the slot scheduler below only needs initialize / reset_slot / track / frame_id / log (tests drive
    # it with a recording stand-in; the product always uses BatchedTracker)
    make = tracker_factory or (lambda: BatchedTracker(network, params, update_intervals=update_intervals, n_mod=n_mod,
                                                      use_template_cache=False))
    trk = make()
    slots = [queue.popleft() for _ in range(B)]
    first = [_load(s.frames[0], reader, n_mod) for s in slots]
    t0 = time.perf_counter()
    trk.initialize(first, [s.init_bbox for s in slots], capacity_hw=capacity_hw)
    init_t = (time.perf_counter() - t0) / B
    pos = [1] * B                                   # next frame index of each slot's sequence
    start = [0] * B                                 # result-table row holding the slot's initial box
    times = [[init_t] for _ in range(B)]
    live = [True] * B
    pool = None
    if prefetch_workers and prefetch_workers > 0:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=int(prefetch_workers), thread_name_prefix="mmt-frames")
    nxt = [None] * B                                # future of the slot's next frame (decoded ahead of the step)

    def fetch(b):
        nxt[b] = None
        if pool is not None and pos[b] < len(slots[b].frames):
            nxt[b] = pool.submit(_load, slots[b].frames[pos[b]], reader, n_mod)

    def take(b):
        f, nxt[b] = nxt[b], None
        return f.result() if f is not None else _load(slots[b].frames[pos[b]], reader, n_mod)

    for b in range(B):
        fetch(b)

    def finish(b):
        s = slots[b]
        # synchronises, once per sequence; an owned copy (the slot's next sequence reuses the table's current row)
        rows = trk.log[start[b]:start[b] + len(s.frames), b].cpu().numpy().copy()
        out[s.name] = rows
        if results_dir is not None:
            save_tracker_output(results_dir, s, rows, times[b])

    while any(live):
        # retire finished slots, refill from the queue
        for b in range(B):
            if live[b] and pos[b] >= len(slots[b].frames):
                finish(b)
                if queue:
                    slots[b] = queue.popleft()
                    t0 = time.perf_counter()
                    trk.reset_slot(b, _load(slots[b].frames[0], reader, n_mod), slots[b].init_bbox)
                    pos[b], start[b], times[b] = 1, trk.frame_id, [time.perf_counter() - t0]
                    fetch(b)
                else:
                    live[b] = False
        # single-frame sequences end right after initialisation
        if not any(live):
            break
        if any(live[b] and pos[b] >= len(slots[b].frames) for b in range(B)):
            continue
        t0 = time.perf_counter()
        frames = [take(b) if live[b] else None for b in range(B)]
        trk.track(frames, active=None if all(live) else live)
        for b in range(B):
            if live[b]:
                pos[b] += 1
                fetch(b)                            # decode the slot's next frame while the GPU runs this step
        dt = (time.perf_counter() - t0) / sum(live)
        for b in range(B):
            if live[b]:
                times[b].append(dt)
    if pool is not None:
        pool.shutdown(wait=True)
    return out
