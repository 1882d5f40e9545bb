#!/usr/bin/env python
"""Times ONE participant's share of the Cornell 1080p/1024-spp frame (rank 0 of WORLD, tile split) on one GPU,
for several values of RC_SLICES: how finely a tile's sample range has to be cut for the grid to have no tail."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from racer_tracer_b200 import harness  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w, h, spp = 1920, 1080, 1024
cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
job = harness.prepare_job(os.path.join(ROOT, "tests", "golden", "scenes", "cornell_box.yml"), cfg, w, h)
r = harness.CudaRenderer([0])
r.set_stream(torch.cuda.current_stream().cuda_stream)
r.upload(job)
acc = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
full = None
shapes = [(0, None), (1, None)] + [(0, k) for k in (1, 2, 3, 6, 12)] + [(1, k) for k in (2, 3, 4, 5, 6, 7, 8)]
for halving, slices in shapes:
    os.environ["RC_SLICE_HALVING"] = str(halving)       # 0: slice lengths S, S-1, .., 1; 1: n/2, n/4, .., last two equal
    if slices is None:
        os.environ.pop("RC_SLICES", None)
    else:
        os.environ["RC_SLICES"] = str(slices)
    p = harness.make_params(w, h, spp, 20, seed=0, specialize=1, rank=0, world=world)
    for _ in range(3):
        r.render_accumulate(p, acc.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r.render_accumulate(p, acc.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    print(f"world {world} halving={halving} RC_SLICES={slices}: {e0.elapsed_time(e1) / 5:.3f} ms per share, {r.stats().kernel_launches} launches", flush=True)
r.close()
