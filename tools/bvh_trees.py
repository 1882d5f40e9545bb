#!/usr/bin/env python
"""Random scene (484 spheres) under different BVHs over the same primitives: the host's median split
(harness._build_bvh, the reference's Node::build), a full-sweep SAH binary tree built here, and the GPU
LBVH (rc_build_lbvh).  Prints samples/s for each and checks that the images agree (the closest hit does
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


def area(lo, hi):
    d = hi - lo
    return 2.0 * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0])


def build_sah(aabb, idx, nodes, order):
    """Pre-order rc_bvh_node list, one primitive per leaf; split = the minimum of
    area(left) * n_left + area(right) * n_right over the three axes, primitives sorted by box centre."""
    me = len(nodes)
    nd = capi.rc_bvh_node()
    nodes.append(nd)
    if len(idx) == 1:
        nd.bmin[:], nd.bmax[:] = aabb[idx[0], :3], aabb[idx[0], 3:]
        nd.left, nd.right = ~len(order), 1
        order.append(idx[0])
        return me
    best = None
    for a in range(3):
        o = sorted(idx, key=lambda i: (aabb[i, a] + aabb[i, 3 + a], i))
        lo = np.minimum.accumulate(aabb[o, :3], axis=0)
        hi = np.maximum.accumulate(aabb[o, 3:], axis=0)
        rlo = np.minimum.accumulate(aabb[o[::-1], :3], axis=0)[::-1]
        rhi = np.maximum.accumulate(aabb[o[::-1], 3:], axis=0)[::-1]
        for k in range(1, len(o)):
            cost = area(lo[k - 1], hi[k - 1]) * k + area(rlo[k], rhi[k]) * (len(o) - k)
            if best is None or cost < best[0]:
                best = (cost, o[:k], o[k:])
    l = build_sah(aabb, best[1], nodes, order)
    r = build_sah(aabb, best[2], nodes, order)
    nd.left, nd.right = l, r
    for a in range(3):
        nd.bmin[a] = min(nodes[l].bmin[a], nodes[r].bmin[a])
        nd.bmax[a] = max(nodes[l].bmax[a], nodes[r].bmax[a])
    return me


def sah_job(job):
    c = job.scene.c
    n = c.n_prims
    aabb = np.ctypeslib.as_array(c.prim_aabb, shape=(n, 6)).copy()
    nodes, order = [], []
    build_sah(aabb, list(range(n)), nodes, order)
    return harness.with_bvh(job, nodes, order)


def main():
    w, h, spp = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1920, 1080, 64)
    cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
    job = harness.prepare_job("random", cfg, w, h)
    r = harness.CudaRenderer([0])
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    acc = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
    p = harness.make_params(w, h, spp, 20, seed=0)
    images = {}
    for name in ("host-median", "sah", "gpu-lbvh"):
        r.upload(sah_job(job) if name == "sah" else job)
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
        print(f"random {w}x{h}x{spp} bvh={name:12s} lib={os.path.basename(capi.LIB_PATH)}: {ms:9.3f} ms  {w * h * spp / ms / 1e6:7.3f} Gsamples/s  "
              f"seg/sample {st.segments / st.samples:.2f}", flush=True)
    ref = images["host-median"]
    for name, img in images.items():
        d = np.abs(img - ref)
        print(f"image {name:12s} vs host-median: max abs diff {d.max():.3e}, pixels differing {(d.reshape(-1, 3).max(axis=1) > 0).mean():.2e}")
    r.close()


if __name__ == "__main__":
    main()
