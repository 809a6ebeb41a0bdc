"""TEST / BASELINE INFRASTRUCTURE - makes the UNMODIFIED reference travel to the GPU box.

`/root/reference` exists only in the build container.  This script copies the files the per-frame forward needs
(lib/, experiments/, external/PreciseRoIPooling/ - 7 MB of Python, YAML and kernel sources) into `baseline/_ref/`,
which is git-ignored (the history stays free of reference sources) but NOT gpurun-ignored, so the snapshot that goes
to the B200 box carries it.  There `oracle/ref_shims.py` imports the reference modules from `baseline/_ref` exactly as
it imports them from `/root/reference` here, and bench.py times them:
  * `bench.py --impl reference`        : the reference forward on the host cores   (cpu_baseline.kind = "reference")
  * `gpu_eager_baseline` of the main arm: the same modules in eager mode on the B200 (fp32 = the reference's stock
                                          path; bf16 autocast beside it)
Nothing under baseline/_ref is imported by the product package; a missing baseline/_ref only downgrades the two
baselines to the pinned CPU port (stated in their `kind`).

Run:  python oracle/ship_ref.py            (also called by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("MMT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ("lib", "experiments", os.path.join("external", "PreciseRoIPooling"), "LICENSE")


def _newest(path):
    m = 0.0
    for d, _, fs in os.walk(path):
        for f in fs:
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def ship(verbose=False) -> str | None:
    """Copy (or refresh) baseline/_ref from the reference tree.  Returns the destination, or None when the reference
    tree is absent (GPU box: the shipped copy is used as it is)."""
    if not os.path.isdir(os.path.join(SRC, "lib", "models")):
        return None
    stamp = os.path.join(DST, ".shipped_from")
    if os.path.exists(stamp) and os.path.getmtime(stamp) >= max(_newest(os.path.join(SRC, p)) if os.path.isdir(os.path.join(SRC, p))
                                                                else os.path.getmtime(os.path.join(SRC, p)) for p in PARTS):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".git", "*.so", "*.o", "build", "_prroi_pooling*")
    for p in PARTS:
        s, d = os.path.join(SRC, p), os.path.join(DST, p)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=ignore)
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    with open(stamp, "w") as f:
        f.write(f"copied unmodified from {SRC} by oracle/ship_ref.py; git-ignored, travels with the gpurun snapshot only\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"{DST}: {n} files")
    return DST


if __name__ == "__main__":
    if ship(verbose=True) is None:
        sys.stderr.write(f"{SRC} not found: nothing shipped\n")
