#!/usr/bin/env python
"""Random scene (484 spheres) under different BVHs over the same primitives: the host's SAH tree
(harness._build_bvh_sah, the default for large scenes), the reference's median split (harness._build_bvh,
Node::build) and the GPU LBVH (rc_build_lbvh).  Prints samples/s for each and checks that the images agree (the closest hit does
not depend on the tree).

    python tools/bvh_trees.py [W H SPP]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from racer_tracer_b200 import capi, harness  # noqa: E402


def median_job(cfg, w, h):
    """The same scene with the reference's median split (harness._build_bvh)."""
    saved = harness.SAH_MIN_OBJECTS
    try:
        harness.SAH_MIN_OBJECTS = 1 << 30
        return harness.prepare_job("random", cfg, w, h)
    finally:
        harness.SAH_MIN_OBJECTS = saved


def main():
    w, h, spp = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1920, 1080, 64)
    cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
    job = harness.prepare_job("random", cfg, w, h)
    r = harness.CudaRenderer([0])
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    acc = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
    images = {}
    for name in ("host-median", "host-sah", "gpu-lbvh", "host-sah/specialised"):
        p = harness.make_params(w, h, spp, 20, seed=0, specialize=2 if name.endswith("specialised") else 0)
        r.upload(median_job(cfg, w, h) if name == "host-median" else job)
        if name == "gpu-lbvh":
            r.build_lbvh()
        for _ in range(2):
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
        images[name] = acc.cpu().numpy().copy()
        print(f"random {w}x{h}x{spp} bvh={name:22s} lib={os.path.basename(capi.LIB_PATH)}: {ms:9.3f} ms  {w * h * spp / ms / 1e6:7.3f} Gsamples/s  "
              f"seg/sample {st.segments / st.samples:.2f} specialised={st.specialized}", flush=True)
    ref = images["host-median"]
    for name, img in images.items():
        d = np.abs(img - ref)
        print(f"image {name:22s} vs host-median: max abs diff {d.max():.3e}, pixels differing {(d.reshape(-1, 3).max(axis=1) > 0).mean():.2e}")
    r.close()


if __name__ == "__main__":
    main()
