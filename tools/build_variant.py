"""Developer helper: link a NAMED experiment build csrc/libmmt_b200_<name>.so = the shipped objects with ONE source recompiled
under extra macros (e.g. `python tools/build_variant.py x1 attention_tc.cu -DMMT_ATTN_EXP=1`).  Loaded by tools/ scripts only,
through MMT_B200_DEV_LIB=<name> (mmt_b200/_lib.py); the product, the tests and bench.py use libmmt_b200.so."""
import importlib.util, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mmt_build", os.path.join(ROOT, "multi-modal-tracking_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build()
obj = os.path.join(b.OBJ, f"{src[:-3]}__{name}.o")
# VARIANT_SRC=<path>: compile that file (e.g. an older revision written to /tmp) in place of csrc/<src>
path = os.environ.get("VARIANT_SRC") or os.path.join(b.CSRC, src)
subprocess.run([b._nvcc(), *b.NVCC_FLAGS, *flags, "-c", path, "-o", obj], check=True)
objs = [os.path.join(b.OBJ, f) for f in sorted(os.listdir(b.OBJ)) if f.endswith(".o") and "__" not in f and f != src[:-3] + ".o"]
out = os.path.join(b.CSRC, f"libmmt_b200_{name}.so")
subprocess.run([b._nvcc(), "-shared", "-o", out, *objs, obj, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
print(out)
