#!/usr/bin/env python
"""Interactive latency of §8(f)-2 on one GPU: what the reference's window thread waits for after a scene edit or a
camera move (src/main.rs:174-190, src/scene_controller/interactive.rs:196-267) — rc_upload_scene (the re-upload
after `bvh.changed()`), rc_set_camera + rc_render_preview (config.preview: 40 spp, depth 10, scale 4) and the full
render at the default 600x600 config — as host wall-clock per call, median of N.  Writes one JSON object.

    python tools/preview_latency.py [out.json]
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racer_tracer_b200 import harness  # noqa: E402

cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
W = H = 600
out = {}
r = harness.CudaRenderer([0])
for name in ("cornell_box", "three_balls", "sandbox", "random"):
    path = name if name == "random" else os.path.join(ROOT, "tests", "golden", "scenes", name + ".yml")
    if name == "sandbox":
        path = "sandbox:" + os.path.join(ROOT, "tests", "golden", "scenes", "cornell_box.yml")
    job = harness.prepare_job(path, cfg, W, H)
    sw, sh = harness.preview_scales(cfg, W, H)
    pv = harness.make_params(W, H, cfg.preview.samples, cfg.preview.max_depth, seed=1)          # precompiled kernel: no NVRTC on the interactive path
    full = harness.make_params(W, H, cfg.render.samples, cfg.render.max_depth, seed=1, specialize=2)
    r.upload(job); r.render_preview(pv, sw, sh); r.render(full)                                  # warm up (+ NVRTC compile of the full render)
    t_up, t_pv, t_full = [], [], []
    for _ in range(15):
        t0 = time.perf_counter(); r.upload(job); t1 = time.perf_counter()
        r.render_preview(pv, sw, sh); t2 = time.perf_counter()
        r.render(full); t3 = time.perf_counter()
        t_up.append(1e3 * (t1 - t0)); t_pv.append(1e3 * (t2 - t1)); t_full.append(1e3 * (t3 - t2))
    out[name] = {"prims": int(job.scene.c.n_prims), "upload_ms": statistics.median(t_up), "set_camera_plus_preview_ms": statistics.median(t_pv),
                 "full_render_600x600_ms": statistics.median(t_full),
                 "preview": f"{cfg.preview.samples} spp, depth {cfg.preview.max_depth}, blocks of {sw}x{sh} pixels",
                 "full": f"{cfg.render.samples} spp, depth {cfg.render.max_depth}, host f64 image (8.6 MB) included"}
    print(name, out[name], flush=True)
r.close()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
