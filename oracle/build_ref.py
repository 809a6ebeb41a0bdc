"""TEST INFRASTRUCTURE - builds the reference's OWN native kernels for this path into oracle/_ref/ (git-ignored, but
shipped to the GPU box with the repo snapshot), from the sources where they lie under /root/reference.  Nothing is
copied; nothing here is on the product path.  The `-m gpu` tests load these libraries when present and compare the
product kernels with them (tests/test_reference_kernels_gpu.py).

  libprroi_ref.so  external/PreciseRoIPooling/src/prroi_pooling_gpu_impl.cu (plain CUDA + extern "C" launchers;
                   PrRoIPoolingForwardGpu is what prroi_pooling_gpu.c:22-44 calls) - compiles unmodified.
  libmsda_ref.so   lib/models/mixformer_vit_rgbt/deformable_attention/ops/src/cuda/ms_deform_im2col_cuda.cuh through
                   oracle/msda_ref_wrapper.cu (+ a one-line THC include shim, oracle/shim_include/).

The reference's own build system (setup.py / JIT load) is not run.  Run:  python oracle/build_ref.py
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("MMT_REFERENCE_ROOT", "/root/reference")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-4000:])
        raise RuntimeError("reference kernel build failed")


def _stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def build(verbose=False):
    """Returns the list of libraries built (or already fresh); [] when the reference tree is absent (GPU box)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "external", "PreciseRoIPooling")):
        return []
    os.makedirs(OUT, exist_ok=True)
    nvcc = _nvcc()
    libs = []
    src = os.path.join(REFERENCE_ROOT, "external", "PreciseRoIPooling", "src", "prroi_pooling_gpu_impl.cu")
    lib = os.path.join(OUT, "libprroi_ref.so")
    if _stale(lib, [src, __file__]):
        _run([nvcc, *ARCH, "-O2", "-shared", "-Xcompiler", "-fPIC", "-w", src, "-o", lib, "-lcudart"])
    libs.append(lib)
    msda_dir = os.path.join(REFERENCE_ROOT, "lib", "models", "mixformer_vit_rgbt", "deformable_attention", "ops", "src", "cuda")
    wrap = os.path.join(HERE, "msda_ref_wrapper.cu")
    lib = os.path.join(OUT, "libmsda_ref.so")
    if _stale(lib, [wrap, os.path.join(msda_dir, "ms_deform_im2col_cuda.cuh"), __file__]):
        import torch
        tinc = os.path.join(os.path.dirname(torch.__file__), "include")
        _run([nvcc, *ARCH, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-w", "-I", os.path.join(HERE, "shim_include"),
              "-I", msda_dir, "-I", tinc, "-I", os.path.join(tinc, "torch", "csrc", "api", "include"), wrap, "-o", lib,
              "-lcudart"])
    libs.append(lib)
    if verbose:
        print("\n".join(libs))
    return libs


if __name__ == "__main__":
    build(verbose=True)
