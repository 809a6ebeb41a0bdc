"""Build recipe for libmmt_b200.so (hand-written sm_100a kernels behind a C ABI).

nvcc cross-compiles for sm_100a without a GPU, so this runs in the CPU container as well as on the
B200 box.  The library is built IN-TREE (csrc/libmmt_b200.so) so that it travels with the repo
snapshot; object files live next to it under csrc/_obj/.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "libmmt_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmmt_b200.so cannot be built")


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return sorted(hs)


def _digest(paths: list[str], extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def source_digest() -> str:
    """Hash of every source + header + flag that goes into the library."""
    return _digest(_sources() + _headers(), " ".join(NVCC_FLAGS))


def is_fresh() -> bool:
    stamp = os.path.join(OBJ, "stamp.txt")
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == source_digest()


def build(verbose: bool = False, force: bool = False, ptxas_info: bool = False, dev: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link csrc/libmmt_b200.so. Returns the library path.
    dev=True builds the DEVELOPER library csrc/libmmt_b200_dev.so instead: the same sources with -DMMT_GEMM_DEV (epilogue
    isolation switches, cycle counters, A/B environment switches) plus csrc/dev/*.cu; selected at import time by
    MMT_B200_DEV_LIB=1 (tools/ only - the product, the tests and bench.py use the shipped library)."""
    if dev:
        return _build_dev(verbose, force, ptxas_info)
    if not force and is_fresh():
        return LIB
    return _compile_and_link(_sources(), list(NVCC_FLAGS), OBJ, LIB, verbose, force, ptxas_info, stamp=True)


def _build_dev(verbose, force, ptxas_info):
    dev_dir = os.path.join(CSRC, "dev")
    srcs = _sources() + sorted(os.path.join(dev_dir, f) for f in os.listdir(dev_dir) if f.endswith(".cu"))
    extra = os.environ.get("MMT_DEV_EXTRA_FLAGS", "").split()        # one-off experiment macros (developer library only)
    return _compile_and_link(srcs, list(NVCC_FLAGS) + ["-DMMT_GEMM_DEV"] + extra, os.path.join(CSRC, "_obj", "dev"),
                             os.path.join(CSRC, "libmmt_b200_dev.so"), verbose, force, ptxas_info, stamp=False)


def _compile_and_link(sources, base_flags, obj_dir, lib_path, verbose, force, ptxas_info, stamp):
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    hdr_digest = _digest(_headers(), " ".join(base_flags))
    flags = list(base_flags) + (["-Xptxas", "-v"] if ptxas_info else [])

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        tag = obj + ".digest"
        want = _digest([src], hdr_digest)
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == want:
            return obj
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0 or ptxas_info:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}")
        with open(tag, "w") as f:
            f.write(want)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"link of {os.path.basename(lib_path)} failed")
    if stamp:
        with open(os.path.join(OBJ, "stamp.txt"), "w") as f:
            f.write(source_digest())
    return lib_path


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_info="--ptxas" in sys.argv, dev="--dev" in sys.argv))
