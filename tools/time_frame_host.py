"""Developer tool: where the host time of BatchedTracker.track() goes (packing memcpy / H2D enqueue / launches)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mmt_b200  # noqa
from mmt_b200 import synthetic, frames as F

variant, B = "mixformer_vit_rgbt", 64
model, cfg = synthetic.make_model(variant, 0)
model = model.cuda()
rng = np.random.default_rng(5)
H, W = 480, 640
sets = [[[rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)] for _ in range(B)] for _ in range(2)]
init = np.stack([rng.uniform(100, 400, B), rng.uniform(100, 300, B), rng.uniform(30, 120, B), rng.uniform(30, 120, B)], 1)
params = types.SimpleNamespace(template_factor=2.0, template_size=128, search_factor=5.0, search_size=288)
trk = F.BatchedTracker(model, params, update_intervals=[10 ** 9], n_mod=2, capacity=256)
trk.initialize(sets[0], init)
for t in range(3):
    trk.track(sets[t & 1])
torch.cuda.synchronize()
print("cpus", os.cpu_count())
# packing alone
imgs = trk._flatten(sets[0])
t0 = time.perf_counter()
for _ in range(10):
    k = trk.up.upload(imgs)
    trk.up.release(k)
torch.cuda.synchronize()
print("upload() incl. H2D wait: %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
host = trk.up.pinned[0].numpy()
t0 = time.perf_counter()
for _ in range(10):
    for im, o in zip(imgs, trk.up.offsets):
        np.copyto(host[o:o + H * W * 3].reshape(H, W, 3), im)
print("serial memcpy: %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
# launches alone (no sync): host time of one forward
t0 = time.perf_counter()
for _ in range(10):
    model(trk._model_args(trk.template), trk._model_args(trk.online_template), trk._model_args(trk.search))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("forward host enqueue: %.2f ms/step (device drain after: %.2f ms)" % ((t1 - t0) / 10 * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter()
for t in range(20):
    trk.track(sets[t & 1])
t1 = time.perf_counter()
torch.cuda.synchronize()
print("track(): host %.2f ms/step, with drain %.2f ms/step" % ((t1 - t0) / 20 * 1e3, (time.perf_counter() - t0) / 20 * 1e3))
