"""The N > 1 host logic on CPU: two gloo ranks partition the image exactly as the
GPU ranks do (rc_partition is pure host arithmetic), exchange their shares with
the same torch.distributed reduce bench.py issues over NCCL, and rank 0 ends up
with every pixel written exactly once (tile split) or every sample counted
exactly once (sample split)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from racer_tracer_b200 import capi, harness


def _worker(rank, world, port, split, w, h, spp, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = harness.make_params(w, h, spp, 20, split=split, rank=rank, world=world)
    sh = harness.partition(p, rank, world)
    mask = harness.share_mask(p, rank, world)
    # stand-in for the traced radiance sums: pixel index + 1, once per traced sample
    n_samples = sh["s_end"] - sh["s_begin"]
    pix = (np.arange(w * h, dtype=np.float64).reshape(h, w) + 1.0)
    accum = torch.from_numpy(np.where(mask, pix * n_samples, 0.0)[..., None].repeat(3, axis=2).copy())
    dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)      # the exchange step of bench.py
    cover = torch.from_numpy(mask.astype(np.int32))
    dist.all_reduce(cover, op=dist.ReduceOp.SUM)
    ranges = [None] * world
    dist.all_gather_object(ranges, (sh["s_begin"], sh["s_end"]))
    if rank == 0:
        np.save(os.path.join(out_dir, "accum.npy"), accum.numpy())
        np.save(os.path.join(out_dir, "cover.npy"), cover.numpy())
        np.save(os.path.join(out_dir, "ranges.npy"), np.array(ranges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("split", [capi.RC_SPLIT_TILES, capi.RC_SPLIT_SAMPLES])
@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_exchange_reassembles_the_image(cuda_lib, tmp_path, split, world):
    w, h, spp = 101, 67, 37      # ragged against the 16x8 tiles and against world
    port = 29500 + (os.getpid() + 7 * world + split) % 2000
    mp.spawn(_worker, args=(world, port, split, w, h, spp, str(tmp_path)), nprocs=world, join=True)
    accum = np.load(tmp_path / "accum.npy")
    cover = np.load(tmp_path / "cover.npy")
    ranges = np.load(tmp_path / "ranges.npy")
    pix = (np.arange(w * h, dtype=np.float64).reshape(h, w) + 1.0)
    assert np.array_equal(accum[..., 0], pix * spp)          # every pixel x sample counted exactly once
    if split == capi.RC_SPLIT_TILES:
        assert (cover == 1).all()                            # disjoint, complete tile cover
        assert all(tuple(r) == (0, spp) for r in ranges)
    else:
        assert (cover == world).all()                        # every rank traces every pixel
        assert ranges[0][0] == 0 and ranges[-1][1] == spp
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_partition_matches_the_documented_interleave(cuda_lib):
    p = harness.make_params(1920, 1080, 1024, 20)
    shares = [harness.partition(p, r, 8) for r in range(8)]
    assert shares[0]["tiles_x"] == 120 and (shares[0]["tile_w"], shares[0]["tile_h"]) == (16, 8)
    assert sum(s["n_tiles"] for s in shares) == 120 * 135
    assert [s["tile_first"] for s in shares] == list(range(8)) and all(s["tile_stride"] == 8 for s in shares)
    q = harness.make_params(3840, 2160, 4096, 20, split=capi.RC_SPLIT_SAMPLES)
    assert [(harness.partition(q, r, 8)["s_begin"], harness.partition(q, r, 8)["s_end"]) for r in range(8)] == \
        [(512 * r, 512 * (r + 1)) for r in range(8)]
    bad = harness.make_params(64, 64, 1, 1)
    with pytest.raises(capi.RacerCudaError):
        harness.partition(bad, 3, 2)
