import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from racer_tracer_b200 import capi, harness
from conftest import scene_path
cfg = harness.load_config("tests/golden/config.yml")
w, h, spp = 333, 201, 24
job = harness.prepare_job(scene_path("three_balls"), cfg, w, h)
r1 = harness.CudaRenderer([0]); r1.upload(job)
one = r1.render(harness.make_params(w, h, spp, 20, seed=4))
r2 = harness.CudaRenderer([0, 1]); r2.upload(job)
for name, kw in [("tiles", dict(split=capi.RC_SPLIT_TILES)), ("samples", dict(split=capi.RC_SPLIT_SAMPLES)), ("wavefront", dict(variant=capi.RC_VARIANT_WAVEFRONT))]:
    two = r2.render(harness.make_params(w, h, spp, 20, seed=4, **kw))
    d = np.abs(one - two).max(axis=2)
    ys, xs = np.nonzero(d > 1e-5)
    print(name, "max", d.max(), "n_bad", len(ys), "first bad", list(zip(ys[:5], xs[:5])), "nan", np.isnan(two).sum(), "zeros", (two.sum(axis=2) == 0).sum())
    if len(ys):
        print("  bad tiles (ty,tx) sample:", sorted(set((int(y)//8, int(x)//16) for y, x in zip(ys, xs)))[:10], "count", len(set((int(y)//8, int(x)//16) for y, x in zip(ys, xs))))
