"""Phase cycle stamps of the attention kernel (developer build with -DMMT_ATTN_EXP=16: MMT_B200_DEV_LIB=ax16)."""
import ctypes, os, sys, subprocess
import numpy as np
os.environ.setdefault("MMT_B200_DEV_LIB", "ax16")
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["sym"])
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_attn.py")).read())
from mmt_b200 import _lib
buf = np.zeros((512, 12), dtype=np.uint64)
f = _lib.lib.mmt_dev_attn_stamps
f.restype = ctypes.c_int
assert f(buf.ctypes.data_as(ctypes.c_void_p)) == 0
b = buf[buf[:, 9] > 0].astype(np.float64)
names = ["wait S", "ld S", "max+exchange", "exp (first half)", "wait P free", "rescale+st+exp 2nd half", "wait O", "epilogue", "blocks", "items", "total"]
tot = np.median(b[:, 10])
print(f"CTAs {len(b)}; median total cycles {tot:.0f}; blocks/CTA {np.median(b[:,8]):.1f}; items/CTA {np.median(b[:,9]):.1f}")
for i in range(8):
    per = b[:, i] / (b[:, 8] if i < 6 else b[:, 9])
    print(f"{names[i]:28s} {np.median(b[:, i]):9.0f} cycles/CTA ({100*np.median(b[:, i])/tot:5.1f} %)   {np.median(per):7.0f} per {'block' if i < 6 else 'item'}")
