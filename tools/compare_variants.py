#!/usr/bin/env python
"""Times both integrator variants (megakernel / wavefront) on every BASELINE scene at a
reduced sample count and prints samples/s, segments/sample and kernel launches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from racer_tracer_b200 import capi, harness  # noqa: E402

CASES = [("three_balls", 600, 600, 200), ("emissive", 600, 600, 200), ("noise_and_textures", 600, 600, 200),
         ("cornell_box", 1920, 1080, 128), ("clown", 3840, 2160, 32), ("sandbox_boxes", 600, 600, 200),
         ("random", 1920, 1080, 32)]
cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
r = harness.CudaRenderer([0])
r.set_stream(torch.cuda.current_stream().cuda_stream)
for name, w, h, spp in CASES:
    job = harness.prepare_job(name if name == "random" else os.path.join(ROOT, "tests", "golden", "scenes", name + ".yml"), cfg, w, h)
    r.upload(job)
    acc = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
    for vname, variant in (("megakernel", capi.RC_VARIANT_MEGAKERNEL), ("wavefront", capi.RC_VARIANT_WAVEFRONT)):
        p = harness.make_params(w, h, spp, 20, seed=0, variant=variant)
        for _ in range(2):
            acc.zero_()
            r.render_accumulate(p, acc.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(3):
            acc.zero_()
            r.render_accumulate(p, acc.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        wall = (time.perf_counter() - t0) / 3 * 1e3
        st = r.stats()
        print(f"{name:20s} {w}x{h}x{spp:4d} {vname:10s} {ms:9.3f} ms (wall {wall:8.3f}) "
              f"{w * h * spp / ms / 1e3:9.1f} Msamples/s  seg/sample {st.segments / st.samples:5.2f} "
              f"launches {st.kernel_launches}", flush=True)
r.close()
