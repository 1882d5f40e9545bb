#!/usr/bin/env python
"""Debug aid: the sample-split part of test_partitions_cover_the_image_exactly with the differences printed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from racer_tracer_b200 import capi, harness

cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
w, h, spp = 333, 201, 12
job = harness.prepare_job(os.path.join(ROOT, "tests", "golden", "scenes", "three_balls.yml"), cfg, w, h)
r = harness.CudaRenderer([0])
r.upload(job)

def acc(p):
    a = torch.zeros(p.height * p.width * 3, dtype=torch.float32, device="cuda:0")
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.render_accumulate(p, a.data_ptr())
    torch.cuda.synchronize()
    return a

whole = acc(harness.make_params(w, h, spp, 20, seed=2))
again = acc(harness.make_params(w, h, spp, 20, seed=2))
print("whole deterministic:", bool(torch.equal(whole, again)))
for world in (2, 3, 8, 12):
    parts = [acc(harness.make_params(w, h, spp, 20, seed=2, rank=k, world=world, split=capi.RC_SPLIT_SAMPLES)) for k in range(world)]
    s = torch.stack(parts).sum(dim=0)
    d = (s - whole).abs().reshape(h, w, 3).amax(dim=2)
    bad = d > 1e-5 * (1 + whole.reshape(h, w, 3).amax(dim=2).abs())
    ys, xs = torch.nonzero(bad, as_tuple=True)
    print(f"world {world}: max diff {float(d.max()):.3e}, pixels off {int(bad.sum())}",
          [(int(y), int(x), float(d[y, x])) for y, x in list(zip(ys.tolist(), xs.tolist()))[:8]])
