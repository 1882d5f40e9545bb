#!/usr/bin/env python
"""Writes tests/golden/vectors/<scene>.npz from the CPU oracle: regression pins for the oracle itself (CPU
tests) and reference outputs the CUDA path is compared with without running the oracle (GPU tests).

The reference (Rust, not buildable here) has no golden vectors of its own: these are outputs
of the restated algorithm under the Philox streams of DESIGN.md §4, regenerated only when that specification
changes.  Per scene: primary-hit AOV at 64x48 with the fixed jitter (ids, t, normals), and a 48x36 image of
4 spp, depth 20, seed 1 (direct sampler), plus the preview renderer's image for cornell_box.

    python tools/make_golden_vectors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from racer_tracer_b200 import harness  # noqa: E402

SCENES = ["three_balls", "emissive", "noise_and_textures", "cornell_box", "clown", "sandbox_boxes", "random", "sandbox"]
OUT = os.path.join(ROOT, "tests", "golden", "vectors")
IMAGES = os.path.join(ROOT, "tests", "golden", "resources", "images")


def scene_path(name):
    if name == "random":
        return name
    if name == "sandbox":
        return "sandbox:" + os.path.join(ROOT, "tests", "golden", "scenes", "cornell_box.yml")
    return os.path.join(ROOT, "tests", "golden", "scenes", name + ".yml")


def vectors(name, cfg):
    path = scene_path(name)
    job = harness.prepare_job(path, cfg, 64, 48, seed=0, image_dirs=[IMAGES])
    ids, t, nrm, _ = O.primary_aov(job, harness.make_params(64, 48, 1, 20, fixed_jitter=1))
    job2 = harness.prepare_job(path, cfg, 48, 36, seed=0, image_dirs=[IMAGES])
    img = O.render(job2, harness.make_params(48, 36, 4, 20, seed=1), threads=1)
    out = {"aov_ids": ids.reshape(48, 64), "aov_t": t.reshape(48, 64), "aov_normal": nrm.reshape(48, 64, 3), "image": img}
    if name == "cornell_box":
        sw, sh = harness.preview_scales(cfg, 48, 36)
        out["preview"] = O.render_preview(job2, harness.make_params(48, 36, 6, 10, seed=1), sw, sh)
        out["preview_scale"] = np.array([sw, sh])
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
    for name in SCENES:
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **vectors(name, cfg))
        print("wrote", name)
