#!/usr/bin/env python
"""Writes tests/golden/reference_renders/<scene>.npy: the reference's own published renders
(/root/reference/assets/<scene>.png, README.md:26-49 — 600x600 RGBA, produced by the reference at its default
200 spp / depth 20, unseeded) reduced to 60x60 block means of the 8-bit colour values, float32 in [0, 1].

They are the only outputs of the reference that exist (SURVEY §4 "de-facto goldens"; it cannot be built here).
Block means remove most of their sampling noise, which makes them usable as a statistical pin for the oracle and
the CUDA path: tests/test_reference_renders.py.  Runs only where /root/reference exists.

For the two scenes with sharp silhouettes (clown, three_balls) the full-resolution image is kept as well, as the
per-pixel sum R + G + B (uint16, <scene>_rgbsum.npz): the Q1 signature test needs single pixels, not block means.

    python tools/make_reference_render_fixtures.py
"""
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/assets"
OUT = os.path.join(ROOT, "tests", "golden", "reference_renders")
SCENES = ["three_balls", "emissive", "noise_and_textures", "clown", "cornell_box"]
N = 60


def block_means(rgb: np.ndarray, n: int = N) -> np.ndarray:
    h, w, _ = rgb.shape
    return rgb.reshape(n, h // n, n, w // n, 3).mean(axis=(1, 3))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name in SCENES:
        im = Image.open(os.path.join(SRC, name + ".png"))
        assert im.size == (600, 600), im.size
        rgb = np.asarray(im.convert("RGB"), dtype=np.float64) / 255.0
        np.save(os.path.join(OUT, name + ".npy"), block_means(rgb).astype(np.float32))
        print("wrote", name, block_means(rgb).mean(axis=(0, 1)).round(4))
        if name in ("clown", "three_balls"):
            rgb8 = np.asarray(im.convert("RGB"), dtype=np.uint16)
            np.savez_compressed(os.path.join(OUT, name + "_rgbsum.npz"), rgbsum=rgb8.sum(axis=2).astype(np.uint16))
