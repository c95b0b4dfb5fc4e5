"""N > 1 host logic on CPU: world_size-2 gloo process group, the oracle injected as the per-rank renderer.
Checks the sample split, the reduce onto rank 0 and that the 2-rank image equals the 1-rank image."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import learn_path_tracing_b200 as L
from learn_path_tracing_b200 import scenes
from conftest import ROOT

W, H, SPP, DEPTH = 48, 27, 13, 8  # odd spp: uneven split 7 + 6


def test_split_samples_partitions_the_range():
    for spp in (0, 1, 7, 13, 256):
        for ws in (1, 2, 3, 8):
            parts = [L.split_samples(spp, ws, r) for r in range(ws)]
            assert sum(c for _, c in parts) == spp
            pos = 0
            for off, cnt in parts:
                assert off == pos
                pos += cnt
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _oracle_accum(world, cam):
    from oracle import ptoracle as O

    def fn(offset, count):
        acc, _, _ = O.render(O.scene_from_world(world), cam.to_struct(), W, H, count, DEPTH, L.PT_SHADE_V2, seed=9,
                             spp_offset=offset, threads=2)
        t = torch.zeros((H * W, 4), dtype=torch.float32)
        t[:, :3] = torch.from_numpy(np.ascontiguousarray(acc.transpose(1, 0, 2))).reshape(H * W, 3)
        return t
    return fn


def _worker(rank, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    world, cam = scenes.scene_8_refract((W, H))
    img = L.render_distributed(world, cam, spp=SPP, propagate_limit=DEPTH, render_accum=_oracle_accum(world, cam),
                               postprocess=False)
    if rank == 0:
        np.save(out_path, img)
    else:
        assert img is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_render_equals_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(port, out), nprocs=2, join=True)
    two = np.load(out)
    world, cam = scenes.scene_8_refract((W, H))
    one = L.render_distributed(world, cam, spp=SPP, propagate_limit=DEPTH, render_accum=_oracle_accum(world, cam),
                               postprocess=False)
    assert two.shape == (W, H, 3)
    # same set of paths (counter-based RNG keyed on the absolute sample index): equal up to fp32 summation order
    assert np.allclose(two, one, rtol=1e-5, atol=1e-6)
    assert two.mean() > 0.1
