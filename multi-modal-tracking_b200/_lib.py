"""ctypes binding of libmmt_b200.so (the C ABI declared in include/mmt_b200.h).

There is deliberately no fallback: if the library is missing or a symbol is absent the import of
this module raises, and every op raises RuntimeError on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMT_B200_DEV_LIB=1 (tools/ only): the developer build with the GEMM experiment switches (build.py --dev)
_DEV = os.environ.get("MMT_B200_DEV_LIB", "")
LIB_PATH = os.path.join(_HERE, "csrc", "libmmt_b200.so" if _DEV in ("", "0") else
                        "libmmt_b200_dev.so" if _DEV == "1" else f"libmmt_b200_{_DEV}.so")   # named experiment builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mmt_b200.h")


class MMTLibraryError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise MMTLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the tracker forward)")
    return ctypes.CDLL(LIB_PATH)


lib = _load()


def declared_symbols() -> list[str]:
    """Every function name declared in include/mmt_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(mmt_[a-z0-9_]+)\s*\(", text)))


def check(status: int, what: str) -> None:
    if status != 0:
        kind = "argument check" if status >= 1000001 else "CUDA"
        raise RuntimeError(f"{what} failed: status {status} ({kind} error)")


def fn(name: str):
    f = getattr(lib, name, None)
    if f is None:
        raise MMTLibraryError(f"symbol {name} missing from {LIB_PATH}")
    f.restype = ctypes.c_int
    return f
