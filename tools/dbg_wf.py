import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racer_tracer_b200 import capi, harness
cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
r = harness.CudaRenderer([0])
for name in ["emissive", "noise_and_textures"]:
    for (w, h, spp) in [(150, 90, 40), (150, 90, 32), (150, 90, 1), (160, 96, 4)]:
        job = harness.prepare_job(os.path.join(ROOT, "tests", "golden", "scenes", name + ".yml"), cfg, w, h)
        r.upload(job)
        a = r.render(harness.make_params(w, h, spp, 20, seed=5, variant=0))
        b = r.render(harness.make_params(w, h, spp, 20, seed=5, variant=1))
        b2 = r.render(harness.make_params(w, h, spp, 20, seed=5, variant=1))
        d = np.abs(a - b).max(axis=2)
        bad = np.argwhere(d > 1e-4)
        print(name, w, h, spp, "max", d.max(), "n bad", len(bad), "wf repeatable", np.array_equal(b, b2), bad[:6].tolist())
        for (y, x) in bad[:3]:
            print("   ", a[y, x], b[y, x])
