import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from racer_tracer_b200 import harness, capi
from oracle import oracle as O
cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
w, h, spp = 300, 300, 32
job = harness.prepare_job(os.path.join(ROOT, "tests", "golden", "scenes", "sandbox_boxes.yml"), cfg, w, h)
r = harness.CudaRenderer([0])
r.upload(job)
imgs = {}
for name, spec in (("pre", 0), ("spec", 1)):
    p = harness.make_params(w, h, spp, 20, seed=6, specialize=spec)
    imgs[name] = r.render(p)
    st = r.stats()
    print(name, st.segments / st.samples)
os.environ["RC_SCENE_MODE"] = "smem"
r.upload(job)
imgs["bvh"] = r.render(harness.make_params(w, h, spp, 20, seed=6))
st = r.stats(); print("bvh", st.segments / st.samples)
ref, cnt = O.render(job, harness.make_params(w, h, spp, 20, seed=6), want_counters=True)
c = cnt.as_dict(); print("oracle", c["segments"] / c["samples"])
for k, v in imgs.items():
    e = np.abs(v - ref).max(axis=2)
    print(k, "frac>2e-3", (e > 2e-3).mean(), "mean img", v.mean(), "ref mean", ref.mean())
