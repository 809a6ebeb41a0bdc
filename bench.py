#!/usr/bin/env python
"""Headline benchmark: tracked frames/s of the per-frame network forward, batched sequences per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one full network forward (both modalities, template + online template + search) for B = 64 sequences
on every GPU: BASELINE.json configs[1] (experiments/mixformer_vit_rgbt two-stream MixViT-B RGB-T, bs=64 per B200).
Sequences are independent, so N GPUs run N x 64 sequences with no data-path collective (weak scaling); one NCCL
all-gather per step brings the [64,4] boxes of every rank together (never inside the forward).

Printed (rank 0, ONE JSON line): `value` = frames/s with inputs resident in HBM; `e2e` = the same through
FrameStep.step() with pinned HOST buffers in the reference tracker loop's data flow (templates resident on the device,
uint8 search crops H2D + device Preprocessor + full forward + D2H of the boxes + synchronise, every step;
`e2e_fp32_crops`: all crops as fp32 host tensors every step); `frame_path` = raw uint8 frames in through BatchedTracker;
`roofline` = the tcgen05 GEMM kernel timed per launch with CUDA events in a second pass of the same K steps, against the
measured cuBLAS bf16 peak of MEASURED_PEAKS.json; `cpu_baseline` = the CPU port of the reference forward (oracle/) on the
host cores, bounded sample.  `--impl reference` times only that CPU port (the reference is Python + torch and
/root/reference does not travel to the GPU box; the port is pinned against the unmodified reference by
oracle/gen_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

VARIANT = "mixformer_vit_rgbt"
BATCH = 64
METRIC = "tracked frames/sec (batched seqs)"
UNIT = "frames/s"
# 2*MAC of matmul + conv of one sequence-frame, reference modules under FlopCounterMode
GFLOP_PER_FRAME = {"mixformer_vit": 92.40, "mixformer_vit_rgbt": 183.79, "mixformer_vit_rgbt_shared": 183.79,
                   "mixformer_vit_rgbt_unibackbone": 183.79, "asymmetric_shared": 186.85,
                   "asymmetric_shared_ce": 143.65, "mixformer_vit_online": 92.45, "mixformer_convmae_online": 118.05}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(bf16_tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm_gbs=d.get("hbm_gbs"),
                    source="measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)")
    return dict(bf16_tflops=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clocks / throttle reasons of this rank's GPU while the timed region runs.  Two sources, both stamped on arrival:
    `nvidia-smi --query-gpu=... -lms 50` started EARLY (construct before the warm-up: on an 8-GPU box nvidia-smi needs longer
    to start than a short timed region lasts) and an NVML polling thread (10 ms).  Only samples that arrived between
    __enter__ and __exit__ count; nvidia-smi rows are preferred, NVML fills in when none arrived in the window."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        self.index, self.rows, self.proc, self.t = index, [], None, None
        self.t0 = self.t1 = None
        self.nvml_rows, self._stop, self.nt, self.h, self.max_mhz = [], threading.Event(), None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(uuid or index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def _poll(self):
        nv = self.nv
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.h)) if reasons_fn else 0
                self.nvml_rows.append((time.perf_counter(), mhz, mask))
            except Exception:
                break
            self._stop.wait(0.01)

    def __enter__(self):
        self.t0 = time.perf_counter()
        if self.h is not None:
            self.nt = threading.Thread(target=self._poll, daemon=True)
            self.nt.start()
        return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()
        self._stop.set()
        if self.nt is not None:
            self.nt.join(timeout=2)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if len(r) < 6 or self.t0 is None or not (self.t0 <= ts <= self.t1):
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                    "source": "nvidia-smi"}
        # NVML bit masks (nvml.h): SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        rows = [r for r in self.nvml_rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        if rows:
            for _, _, mask in rows:
                for n, bit in bits.items():
                    if mask & bit:
                        reasons.add(n)
            return {"sm_mhz": statistics.median([r[1] for r in rows]), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                    "samples": len(rows), "source": "nvml (no nvidia-smi row arrived inside the timed region)"}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}


def cpu_baseline_frames_per_s(variant, yaml_name, budget_s, batch=8, threads=None):
    """Time the reference forward on `threads` host threads for about `budget_s`: the UNMODIFIED reference modules
    when a reference tree is present (build container, or baseline/_ref on the GPU box) - kind "reference" - else
    oracle.forward, the pinned CPU port - kind "port".  The CPU run happens in a CHILD process (the .cuda() calls in the
    reference head's constructor must be neutralised for a CPU model, which cannot be done in the bench process)."""
    threads = threads or os.cpu_count() or 1
    code = (
        "import sys, os, time, json, statistics, torch\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "import mmt_b200\n"
        "from mmt_b200 import synthetic\n"
        f"variant, yaml_name, batch, budget, threads = {variant!r}, {yaml_name!r}, {batch}, {budget_s}, {threads}\n"
        "torch.set_num_threads(threads)\n"
        "model, cfg = synthetic.make_model(variant, 0, yaml_name=yaml_name)\n"
        "sd = model.state_dict()\n"
        "ref, where = bench.build_reference(variant, yaml_name, sd, cpu_model=True)\n"
        "if ref is not None:\n"
        "    kind = 'reference'\n"
        "    def fwd(i):\n"
        "        with torch.no_grad():\n"
        "            return ref(*i)\n"
        "else:\n"
        "    from oracle import mixformer_oracle as O\n"
        "    kind, where = 'port', 'oracle/mixformer_oracle.py'\n"
        "    fwd = lambda i: O.forward(variant, sd, cfg, *i)\n"
        "inputs = synthetic.make_inputs(variant, cfg, batch, 1)\n"
        "fwd(inputs)\n"
        "times, t_end = [], time.perf_counter() + budget\n"
        "while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):\n"
        "    t0 = time.perf_counter(); fwd(inputs); times.append(time.perf_counter() - t0)\n"
        "print('CPUBASE ' + json.dumps({'value': batch / statistics.median(times), 'kind': kind, 'where': where, 'n': len(times)}))\n")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("CPUBASE ")]
    if r.returncode != 0 or not line:
        raise RuntimeError("cpu baseline child failed: " + r.stderr[-2000:])
    d = json.loads(line[-1][8:])
    impl = (f"UNMODIFIED reference modules ({d['where']})" if d["kind"] == "reference"
            else "CPU port of the reference forward (oracle/mixformer_oracle.py)")
    return d["value"], threads, d["kind"], (f"{d['n']} forwards of {batch} sequence-frames of the bench workload, {impl}, fp32, "
                                            f"torch CPU, {threads} threads, median")


def frame_path_bench(model, cfg, variant, B, steps, dev, cpu_budget):
    """Frames/s through BatchedTracker: raw uint8 640x480 frames on the HOST in, tracker state on the
    device out; crop + resize + colour map + normalisation + forward + box map-back/clip per step, one H2D of the
    frames, no synchronisation inside the loop.  Beside it the same frame-side work through the CPU oracle
    (cv2-equivalent arithmetic, one host thread) and the crop kernel's HBM roofline."""
    import types
    import numpy as np
    from mmt_b200 import frames as F, ops
    n_mod = 1 if variant in ("mixformer_vit", "mixformer_vit_online", "mixformer_convmae_online") else 2
    H, W = 480, 640
    rng = np.random.default_rng(5)
    sets = [[[rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(n_mod)] if n_mod > 1
             else rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(B)] for _ in range(2)]
    init = np.stack([rng.uniform(100, 400, B), rng.uniform(100, 300, B), rng.uniform(30, 120, B), rng.uniform(30, 120, B)], 1)
    params = types.SimpleNamespace(template_factor=float(cfg.TEST.TEMPLATE_FACTOR), template_size=int(cfg.TEST.TEMPLATE_SIZE),
                                   search_factor=float(cfg.TEST.SEARCH_FACTOR), search_size=int(cfg.TEST.SEARCH_SIZE))
    # the two-stream tracker class feeds both modalities through Preprocessor_wo_mask (no colour map), the others apply
    # JET to the infrared crop (Preprocessor_Multimodal)
    trk = F.BatchedTracker(model, params, update_intervals=[10 ** 9], n_mod=n_mod, capacity=steps + 8,
                           jet_mask=0 if variant == "mixformer_vit_rgbt" else None)
    trk.initialize(sets[0], init)
    for t in range(3):
        trk.track(sets[t & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for t in range(steps):
        trk.track(sets[t & 1])
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    boxes = trk.results()
    assert np.isfinite(boxes).all()
    # crop kernel alone, CUDA events, search-size crops of the current states
    k = 0
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    c0.record()
    for _ in range(reps):
        trk._crop(k, params.search_factor, params.search_size, trk.search, rf=trk.rf)
    c1.record()
    torch.cuda.synchronize()
    crop_us = c0.elapsed_time(c1) * 1e3 / reps
    st = trk.state.cpu().numpy()
    side = np.ceil(np.sqrt(st[:, 2] * st[:, 3]) * params.search_factor)
    S = params.search_size
    # algorithmic bytes: every output value written once (fp32 CHW) + the source window read once (uint8, clipped to
    # the frame is ignored: an upper bound on the read side)
    alg_bytes = n_mod * float(np.sum(3 * S * S * 4 + np.minimum(side, max(H, W)) ** 2 * 3))
    peaks = _peaks()
    out = {"value": B * steps / max(wall, e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
           "ms_per_step": max(wall, e0.elapsed_time(e1) * 1e-3) * 1e3 / steps, "frame_hw": [H, W], "modalities": n_mod,
           "h2d_bytes_per_step": trk.up.h2d_bytes, "d2h_bytes_per_step": 0,
           "note": "BatchedTracker.track(): uint8 frames on the host -> one H2D -> crop/resize/colour-map/normalise kernel -> "
                   "forward -> box map-back + clip kernel; state and result table stay on the device (no per-frame sync)",
           "crop_kernel": {"bound": "hbm", "avg_launch_us": crop_us, "achieved": alg_bytes / (crop_us * 1e-6) / 1e9,
                           "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": alg_bytes / (crop_us * 1e-6) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg_bytes}}
    if cpu_budget > 0:
        from oracle import frame_oracle as FO
        n, t_end, t0 = 0, time.perf_counter() + min(cpu_budget, 5.0), time.perf_counter()
        while time.perf_counter() < t_end or n < 2:
            b = n % B
            for m in range(n_mod):
                im = sets[0][b][m] if n_mod > 1 else sets[0][b]
                c, _ = FO.sample_target(im, list(init[b]), params.search_factor, params.search_size)
                FO.normalize(FO.apply_jet(c) if (m == 1 and variant != "mixformer_vit_rgbt") else c)
            n += 1
        out["cpu_frame_side"] = {"value": n / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                                 "sample": f"{n} sequence-frames of sample_target + Preprocessor through oracle/frame_oracle.py "
                                           "(numpy restatement of the cv2 arithmetic), frame-side work only, no forward"}
    return out


def workload_config(variant, yaml_name, cfg, B, world, rgbt):
    """The `config` object of the JSON line - identical in the b200 arm and the reference arm."""
    from mmt_b200 import synthetic
    return {"workload": f"{variant} ({yaml_name or synthetic.DEFAULT_YAML[variant]}.yaml) "
                        f"{'RGB-T two-modality' if rgbt else 'RGB'} full forward "
                        f"({cfg.DATA.TEMPLATE.SIZE}^2 template + online template, {cfg.DATA.SEARCH.SIZE}^2 search), "
                        f"bs={B} sequences per GPU, seeded random-init weights, N(0,1) crops",
            "batch_per_gpu": B, "sequences": B * max(1, world), "parallelism": f"sequence-sharded x{max(1, world)}",
            "l2": "working set per step (weights 2x209 MB + >1 GB activations) exceeds the 126 MB L2; no flush needed"}


def build_reference(variant, yaml_name, sd, cpu_model):
    """The UNMODIFIED reference module for `variant` (imported from /root/reference or the shipped copy baseline/_ref
    through oracle/ref_shims.py) carrying the state_dict `sd`; None when no reference tree is available."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        return None, None
    import io
    import contextlib
    import warnings
    from mmt_b200 import synthetic
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        model, _ = ref_shims.build_reference_model(variant, yaml_name or synthetic.DEFAULT_YAML[variant], cpu_model=cpu_model)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model, ref_shims.reference_kind()


def run_reference_arm(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores - the UNMODIFIED
    reference modules when a reference tree is present (/root/reference, or the copy oracle/ship_ref.py ships under
    baseline/_ref), else the pinned CPU port (oracle/).  Each step = one forward of a bounded sample of the bs=64
    workload (sized from a calibration forward so that the whole run stays within a few minutes)."""
    if rank != 0:
        return
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    variant, B = args.variant, args.batch
    model, cfg = synthetic.make_model(variant, 0, yaml_name=args.yaml)
    sd = model.state_dict()
    rgbt = variant not in ("mixformer_vit", "mixformer_vit_online", "mixformer_convmae_online")
    ref, where = build_reference(variant, args.yaml, sd, cpu_model=True)
    if ref is not None:
        kind = "reference"

        def fwd(inputs):
            with torch.no_grad():
                return ref(*inputs)[1]
        what = f"UNMODIFIED reference modules ({where}) through oracle/ref_shims.py, fp32, torch CPU"
    else:
        from oracle import mixformer_oracle as O
        kind = "port"
        fwd = lambda inputs: O.forward(variant, sd, cfg, *inputs)["pred_boxes"]
        what = "CPU port of the reference forward (oracle/mixformer_oracle.py, pinned against the unmodified reference)"
    steps, warmup = args.steps, max(1, args.warmup)
    # calibration: one forward of 4 sequence-frames -> sample size such that (steps + warmup) forwards take ~120 s
    cal = synthetic.make_inputs(variant, cfg, 4, 1)
    fwd(cal)
    t0 = time.perf_counter()
    fwd(cal)
    per_frame = (time.perf_counter() - t0) / 4
    sample = int(max(2, min(B, 120.0 / ((steps + warmup) * per_frame))))
    inputs = synthetic.make_inputs(variant, cfg, sample, 1)
    for _ in range(warmup):
        fwd(inputs)
    t0 = time.perf_counter()
    for _ in range(steps):
        boxes = fwd(inputs)
    dt = (time.perf_counter() - t0) / steps
    assert bool(torch.isfinite(boxes).all())
    v = sample / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(variant, args.yaml, cfg, B, args.gpus, rgbt),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{steps} steps x {sample} sequence-frames of the bs={B} workload per step; {what}, "
                                   f"{threads} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_eager_baseline(variant, yaml_name, sd, dev, dev_inputs, dev1, steps):
    """The UNMODIFIED reference modules in eager mode on this B200 (SURVEY 2b / BASELINE.md section 3 "GPU reference
    beside it"): fp32 as the reference runs it (its trackers never enable autocast or TF32), fp32 with TF32 matmuls
    allowed, and under torch.autocast(bfloat16); at the bench batch and at bs=1.  Deformable attention runs through the
    reference's own CUDA kernel compiled from its sources (oracle/_ref/libmsda_ref.so), else its pure-PyTorch core.
    CUDA-event timing after warm-up; None when no reference tree travelled."""
    ref, where = build_reference(variant, yaml_name, sd, cpu_model=False)
    if ref is None:
        return {"unavailable": "no reference tree on this box (baseline/_ref not shipped)"}
    from oracle import ref_shims
    ref = ref.to(dev)
    out = {"what": f"UNMODIFIED reference modules ({where}), PyTorch eager on the same GPU, same weights and inputs",
           "modes": {}}
    tf32_0 = torch.backends.cuda.matmul.allow_tf32

    def timed(inputs, n, ctx):
        with torch.no_grad(), ctx():
            for _ in range(2):
                ref(*inputs)
            torch.cuda.synchronize()
            evs = []
            for _ in range(n):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                boxes = ref(*inputs)[1]
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in evs), boxes

    import contextlib
    B = (dev_inputs[2][0] if isinstance(dev_inputs[2], list) else dev_inputs[2]).shape[0]
    try:
        for mode in ("fp32", "tf32", "bf16_autocast"):
            torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
            ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode == "bf16_autocast" else contextlib.nullcontext
            ms, boxes = timed(dev_inputs, max(3, min(steps, 8)), ctx)
            rec = {"frames_per_s": B * 1e3 / ms[len(ms) // 2], "ms_per_step": ms[len(ms) // 2], "batch": B}
            if dev1 is not None:
                ms1, _ = timed(dev1, 30, ctx)
                rec["latency_bs1_p50_ms"] = ms1[len(ms1) // 2]
            rec["boxes"] = boxes.detach().float().view(-1, 4)
            out["modes"][mode] = rec
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32_0
    out["msda"] = ref_shims.MSDA_BACKEND["last"]
    del ref
    torch.cuda.empty_cache()
    return out


# The other BASELINE.json configurations as short lines under `configs` of the default run (config 2 is the headline):
# (key, variant, yaml, sequences per GPU, mode, GFLOP per sequence-frame - SURVEY.md section 8d / BASELINE.md section 2)
CONFIG_LINES = [
    ("config3_shared_backbone", "mixformer_vit_rgbt_shared", None, 64, "full", 183.79),
    ("config3_unibackbone", "mixformer_vit_rgbt_unibackbone", None, 64, "full", 183.79),
    ("config4_candidate_elimination_bs128", "asymmetric_shared_ce", None, 128, "full", 143.65),
    ("config4_candidate_elimination_bs128_templates_cached", "asymmetric_shared_ce", None, 128, "tcache", None),
    ("config5_convmae_large_spm_full", "mixformer_convmae_online", "baseline_large", 32, "full", 681.93),
    ("config5_convmae_large_spm_cached_1p3", "mixformer_convmae_online", "baseline_large", 32, "cached", 483.85),
    ("config5_mixvit_large_spm_full", "mixformer_vit_online", "baseline_large", 32, "full", 600.00),
    ("config5_mixvit_large_spm_cached_1p3", "mixformer_vit_online", "baseline_large", 32, "cached", 433.75),
]


def config_lines(dev, steps, precision, peaks):
    """5-step lines of BASELINE.json configs 3, 4 and 5 on this GPU (same timing rules as the headline: >= 3 warm-up
    steps, CUDA events, inputs resident, working sets far beyond L2).  "cached" = the online trackers' real per-frame
    path: set_online_batch() once (1 template + 3 online templates per sequence), forward_test_batch() per frame."""
    import mmt_b200  # noqa: F401
    from mmt_b200 import synthetic
    out = {}
    model, last = None, None
    for key, variant, yaml_name, B, mode, gf in CONFIG_LINES:
        try:
            if last != (variant, yaml_name):
                del model
                torch.cuda.empty_cache()
                model, cfg = synthetic.make_model(variant, 0, yaml_name=yaml_name)
                model = model.to(dev).set_precision(precision)
                last = (variant, yaml_name)
            to_dev = lambda a: [x.to(dev) for x in a] if isinstance(a, (list, tuple)) else a.to(dev)
            t, ot, s_ = [to_dev(a) for a in synthetic.make_inputs(variant, cfg, B, 1)]
            if mode == "tcache":      # template-side reuse: search tokens only per frame, bit-identical boxes
                model.cache_templates(t, ot)
                fn = lambda: model.forward_search(s_)[1]
            elif mode == "cached":
                g = torch.Generator().manual_seed(3)
                ots = torch.randn(B, 3, 3, cfg.DATA.TEMPLATE.SIZE, cfg.DATA.TEMPLATE.SIZE, generator=g).to(dev)
                model.set_online_batch(t, ots)
                fn = lambda: model.forward_test_batch(s_)[1]
            else:
                fn = lambda: model(t, ot, s_)[1]
            for _ in range(3):
                boxes = fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                boxes = fn()
            b.record()
            torch.cuda.synchronize()
            assert bool(torch.isfinite(boxes).all())
            ms = a.elapsed_time(b) / steps
            fps = B * 1e3 / ms
            out[key] = {"variant": variant, "yaml": yaml_name or synthetic.DEFAULT_YAML[variant], "batch": B, "mode": mode,
                        "value": fps, "unit": UNIT, "ms_per_step": ms, "steps": steps, "gflop_per_frame": gf}
            if gf:
                out[key].update(step_tensor_tflops=gf * fps / 1e3, frac_of_measured_peak=gf * fps / 1e3 / peaks["bf16_tflops"],
                                frac_of_nominal_2250=gf * fps / 1e3 / 2250.0)
        except Exception as e:      # a failing side line must not take the headline down; it is reported as such
            out[key] = {"error": f"{type(e).__name__}: {e}"}
    del model
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default=VARIANT)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--yaml", default=None, help="experiment file name of the variant (default: the variant's headline YAML)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline work (rank 0, N=1)")
    ap.add_argument("--no-profile", action="store_true", help="skip per-GEMM CUDA-event bracketing")
    ap.add_argument("--no-latency", action="store_true", help="skip the bs=1 per-frame latency measurement")
    ap.add_argument("--graph", action="store_true", help="developer A/B: replay the bs=B forward as one CUDA graph in the "
                                                        "resident-input region")
    ap.add_argument("--no-frame-path", action="store_true", help="skip the BatchedTracker (uint8 frames in) measurement")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-GPU run of the unmodified reference modules")
    ap.add_argument("--no-variants", action="store_true", help="skip the short lines of BASELINE.json configs 3, 4, 5")
    ap.add_argument("--breakdown", action="store_true",
                    help="developer aid: after the timed regions, run 3 more steps with EVERY op bracketed by CUDA "
                         "events and print the per-class table to stderr")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the forward has no CPU fallback")

    import torch.distributed as dist
    if world > 1:
        # one slice of the host cores per rank: the launch threads of the ranks do not migrate onto each other
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    import mmt_b200  # noqa: F401
    from mmt_b200 import ops, runner, synthetic

    variant, B = args.variant, args.batch
    model, cfg = synthetic.make_model(variant, 0, yaml_name=args.yaml)
    model = model.cuda(local_rank).set_precision(args.precision)
    # device-resident inputs for `value`; pinned host copies for `e2e`
    host_inputs = synthetic.make_inputs(variant, cfg, B, 1 + rank, pin=True)
    to_dev = lambda a: [x.to(dev) for x in a] if isinstance(a, (list, tuple)) else a.to(dev)
    dev_inputs = [to_dev(a) for a in host_inputs]
    n_seq_total = B * world
    owned = runner.shard_sequences(n_seq_total, world, rank)
    assert len(owned) == B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # boxes of every step stay on the device in a [K, B, 4] log; ONE all-gather at the end of the timed region brings
    # the ranks' logs together (the design's only collective: never inside the forward, and not once per step)
    box_log = torch.empty((max(args.steps, 8), B, 4), device=dev, dtype=torch.float32)

    def step_resident(k=0):
        out, coords = model(*dev_inputs)
        box_log[k].copy_(coords.view(-1, 4))
        return coords

    if args.graph:
        model.enable_cuda_graph(True)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    clock_sampler = ClockSampler(local_rank, gpu_uuid)        # started before the warm-up: see the class
    # L2: one step touches >= 2 x 104.7 M bf16 weights (two streams) + ~1 GB of activations, far beyond the 126 MB L2
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()

    ops.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clock_sampler as clocks:
        barrier()
        e0.record()
        for k in range(args.steps):
            step_resident(k)
        gathered_log = runner.gather_boxes(box_log[:args.steps].reshape(-1, 4))      # [world, K*B, 4]
        e1.record()
        barrier()
    launches = ops.LAUNCHES
    gathered = gathered_log.view(gathered_log.shape[0], args.steps, B, 4)[:, -1]
    if args.graph:
        model.enable_cuda_graph(False)
    # ---- roofline pass: the same K steps again with every tensor-core GEMM launch bracketed by CUDA events on its
    # stream.  The two-backbone variant runs its modality chains on two concurrent streams in the headline region, where
    # brackets of one chain would include the other chain's kernels: this pass runs the chains back to back on one stream
    # (same kernels, same data, bit-identical boxes) and reports its own step time beside the kernel numbers.
    prof, prof_ms = None, None
    if not args.no_profile:
        eng = model.engine()
        two = getattr(eng, "_two_streams", None) if variant == "mixformer_vit_rgbt" else None
        if variant == "mixformer_vit_rgbt":
            eng.set_two_streams(False)
        step_resident()
        prof = ops.LaunchProfiler()
        ops.PROFILER = prof
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        p0.record()
        for k in range(args.steps):
            step_resident(k)
        p1.record()
        barrier()
        ops.PROFILER = None
        prof_ms = p0.elapsed_time(p1)
        if variant == "mixformer_vit_rgbt":
            eng.set_two_streams(True if two is None else two)
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_seq_total * args.steps / (ms_max * 1e-3)
    boxes_all = runner.unshard_boxes(gathered, n_seq_total)
    assert boxes_all.shape == (n_seq_total, 4) and bool(torch.isfinite(boxes_all).all())

    # ---- e2e through the public FrameStep API, host buffers in, host boxes out, every step synchronised.
    # (1) `e2e`: the data flow of the reference's tracker loop (lib/test/tracker/asymmetric_shared_ce.py:74-103): the
    #     (online) templates live on the device between frames, every frame uploads the uint8 search crops that
    #     sample_target produced (`torch.tensor(img_arr).cuda()`), normalises them on the device (Preprocessor), runs the
    #     FULL forward (templates + search) and reads the boxes back.
    # (2) `e2e_fp32_crops`: all three crops of every sequence as normalised fp32 host tensors, every step (178 MB).
    def time_e2e(step_fn):
        for _ in range(3):
            step_fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            hb_ = step_fn()
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        assert bool(torch.isfinite(hb_).all())
        return n_seq_total * args.steps / (float(tt.item()) * 1e-3)

    fs = runner.FrameStep(model, dev)
    e2e_fp32 = time_e2e(lambda: fs.step(*host_inputs))
    fp32_h2d = fs.h2d_bytes
    gu = torch.Generator().manual_seed(7 + rank)
    u8 = lambda size: torch.randint(0, 256, (B, size, size, 3), generator=gu, dtype=torch.uint8).pin_memory()
    ts_, ss_ = int(cfg.DATA.TEMPLATE.SIZE), int(cfg.DATA.SEARCH.SIZE)
    rgbt_in = isinstance(dev_inputs[2], list)
    mk = (lambda size: [u8(size), u8(size)]) if rgbt_in else u8
    fs8 = runner.FrameStep(model, dev, jet_mask=0 if variant == "mixformer_vit_rgbt" else None)
    fs8.set_templates(mk(ts_), mk(ts_))
    search_u8 = mk(ss_)
    e2e_value = time_e2e(lambda: fs8.step(None, None, search_u8))
    fs_h2d, fs_d2h = fs8.h2d_bytes, fs8.d2h_bytes

    # ---- template-side reuse (reported separately, NOT the headline: the metric's frame is a full
    # forward): templates cached once, every step runs the search tokens only - bit-identical boxes (tests)
    cached = None
    if hasattr(model, "cache_templates") and variant in ("mixformer_vit", "mixformer_vit_rgbt", "mixformer_vit_rgbt_shared",
                                                         "mixformer_vit_rgbt_unibackbone", "asymmetric_shared",
                                                         "asymmetric_shared_ce"):
        model.cache_templates(dev_inputs[0], dev_inputs[1])
        for _ in range(3):
            model.forward_search(dev_inputs[2])
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for k in range(args.steps):
            _, cb = model.forward_search(dev_inputs[2])
            box_log[k].copy_(cb.view(-1, 4))
        runner.gather_boxes(box_log[:args.steps].reshape(-1, 4))
        c1.record()
        barrier()
        tc = torch.tensor([c0.elapsed_time(c1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        cached = {"value": n_seq_total * args.steps / (float(tc.item()) * 1e-3), "unit": UNIT,
                  "ms_per_step": float(tc.item()) / args.steps,
                  "note": "cache_templates() once + forward_search() per frame: search tokens only against the cached "
                          "per-layer template q/k/v; boxes bit-identical to the full forward"}

    # ---- bs=1 per-frame latency (second half of BASELINE.json's metric), whole forward replayed as one CUDA graph
    lat = None
    if world == 1 and not args.no_latency:
        host1 = synthetic.make_inputs(variant, cfg, 1, 99, pin=True)
        dev1 = [to_dev(a) for a in host1]
        model.enable_cuda_graph(True)
        for _ in range(10):
            model(*dev1)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
        for a, b in evs:
            a.record()
            model(*dev1)
            b.record()
            torch.cuda.synchronize()
        dev_ms = sorted(a.elapsed_time(b) for a, b in evs)
        fs1 = runner.FrameStep(model, dev)
        for _ in range(10):
            fs1.step(*host1)
        wall = []
        for _ in range(200):
            t0 = time.perf_counter()
            fs1.step(*host1)
            wall.append((time.perf_counter() - t0) * 1e3)
        wall.sort()
        model.enable_cuda_graph(False)
        lat = {"batch": 1, "mode": "one CUDA graph replay per frame", "device_p50_ms": dev_ms[100], "device_p99_ms": dev_ms[197],
               "e2e_host_p50_ms": wall[100], "e2e_host_p99_ms": wall[197],
               "note": "device = CUDA events around model(crops on device); e2e_host = wall clock of FrameStep.step "
                       "(pinned host crops -> H2D -> graph replay -> D2H box -> sync)"}

    eager = None
    if world == 1 and rank == 0 and not args.no_eager:
        try:
            eager = gpu_eager_baseline(variant, args.yaml, model.state_dict(), dev, dev_inputs,
                                       [to_dev(a) for a in synthetic.make_inputs(variant, cfg, 1, 99)], args.steps)
            mine = model(*dev_inputs)[1].view(-1, 4).float()
            for rec in eager.get("modes", {}).values():       # same-run parity: our bf16 boxes vs the reference's on this GPU
                rec["max_box_diff_px_vs_b200_path"] = float((rec.pop("boxes") - mine).abs().max()) * int(cfg.DATA.SEARCH.SIZE)
        except Exception as e:
            eager = {"error": f"{type(e).__name__}: {e}"}
    configs = None
    if world == 1 and rank == 0 and not args.no_variants and variant == VARIANT and args.yaml is None:
        configs = config_lines(dev, 5, args.precision, _peaks())

    frame_path = None
    if world == 1 and not args.no_frame_path and variant not in ("mixformer_vit_online", "mixformer_convmae_online"):
        frame_path = frame_path_bench(model, cfg, variant, B, args.steps, dev, args.cpu_budget)

    if args.breakdown and rank == 0:
        bp = ops.LaunchProfiler(all_ops=True)
        ops.PROFILER = bp
        for _ in range(3):
            step_resident()
        torch.cuda.synchronize()
        ops.PROFILER = None
        tab = bp.table()
        tot = sum(v["ms"] for v in tab.values())
        for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"]):
            sys.stderr.write(f"  {k:22s} {v['ms'] / 3:8.3f} ms/step  {v['launches'] // 3:4d} launches  "
                             f"{100 * v['ms'] / tot:5.1f}%\n")
        sys.stderr.write(f"  sum of bracketed ops   {tot / 3:8.3f} ms/step\n")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    roof = None
    if prof is not None:
        # dominant kernel = the CTA-pair instantiation gemm_bf16_tcgen05_kernel<256, true> (the backbone's qkv / proj /
        # fc1 / fc2 and the big fusion GEMMs); the whole GEMM class (+ single-CTA and implicit-conv instantiations) beside it
        n_l, gemm_ms, gemm_flops = prof.summary("gemm_bf16_pair")
        if n_l == 0:
            n_l, gemm_ms, gemm_flops = prof.summary("gemm_bf16")
        a_l, a_ms, a_flops = prof.summary("gemm_bf16")
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
        traffic = traffic_source = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("variant") == variant and tj.get("batch") == B:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_source = tj.get("source")
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel<256, PAIR> (cta_group::2; every launch of it in the timed region)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_source, "peak_source": peaks["source"],
                "launches": n_l, "avg_launch_us": gemm_ms * 1e3 / max(1, n_l),
                "share_of_step": gemm_ms / prof_ms, "gflop_per_launch_avg": gemm_flops / max(1, n_l) / 1e9,
                "measured_over": f"{args.steps} steps right after the timed region, modality chains on ONE stream "
                                 f"({prof_ms / args.steps:.2f} ms/step; the headline region overlaps them on two streams)"
                                 if variant == "mixformer_vit_rgbt" else f"{args.steps} steps right after the timed region",
                "all_gemm_instantiations": {"achieved": a_flops / (a_ms * 1e-3) / 1e12,
                                            "frac": a_flops / (a_ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
                                            "launches": a_l, "share_of_step": a_ms / prof_ms}}
    gf = {("mixformer_vit_online", "baseline_large"): 600.00, ("mixformer_convmae_online", "baseline_large"): 681.93}.get(
        (variant, args.yaml), GFLOP_PER_FRAME.get(variant, 0.0))
    step_tflops = gf * value / world / 1e3
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
        "config": workload_config(variant, args.yaml, cfg, B, world, isinstance(dev_inputs[2], list)),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": fs_h2d, "d2h_bytes_per_step": fs_d2h,
                "note": "FrameStep.step(None, None, search): templates resident on the device (as in the reference tracker "
                        "loop), per step pinned uint8 search crops -> H2D -> device Preprocessor -> FULL forward -> boxes "
                        "D2H -> synchronise"},
        "e2e_fp32_crops": {"value": e2e_fp32, "unit": UNIT, "h2d_bytes_per_step": fp32_h2d, "d2h_bytes_per_step": fs_d2h,
                           "note": "FrameStep.step(template, online_template, search) with normalised fp32 host crops, all "
                                   "three uploaded every step"},
        "gpu_launches": launches,
        "roofline": roof,
        "step_tensor_tflops_per_gpu": step_tflops,
        "step_tensor_frac_of_measured_peak": step_tflops / peaks["bf16_tflops"],
        "step_tensor_frac_of_nominal_2250": step_tflops / 2250.0,
        "clocks": clocks.summary(),
        "latency_bs1": lat,
        "cached_template": cached,
        "frame_path": frame_path,
        "gpu_eager_baseline": eager,
        "configs": configs,
    }
    if world == 1 and args.cpu_budget > 0:
        v, cores, kind, sample = cpu_baseline_frames_per_s(variant, args.yaml, args.cpu_budget)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
