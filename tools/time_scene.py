#!/usr/bin/env python
"""Times one scene with the precompiled and the scene-specialised megakernel:  time_scene.py SCENE W H SPP"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from racer_tracer_b200 import harness  # noqa: E402

name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
job = harness.prepare_job(name if name == "random" else os.path.join(ROOT, "tests", "golden", "scenes", name + ".yml"), cfg, w, h)
r = harness.CudaRenderer([0])
r.set_stream(torch.cuda.current_stream().cuda_stream)
r.upload(job)
acc = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
for spec in (0, 2):
    p = harness.make_params(w, h, spp, 20, seed=0, specialize=spec)
    for _ in range(3):
        acc.zero_()
        r.render_accumulate(p, acc.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        acc.zero_()
        r.render_accumulate(p, acc.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    st = r.stats()
    print(f"{name} {w}x{h}x{spp} specialize={spec}: {ms:.3f} ms  {w * h * spp / ms / 1e6:.2f} Gsamples/s  seg/sample {st.segments / st.samples:.2f}", flush=True)
r.close()
