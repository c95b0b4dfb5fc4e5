"""N-GPU image == 1-GPU image on the CUDA path (BASELINE configs[3]: sample ranges split over the GPUs,
accumulators summed by a reduce), for a v2 scene AND a legacy mesh scene, through the public API
(render_distributed, LegacyRenderer) — not through an injected renderer like tests/test_multigpu_gloo.py.

Two ranks are spawned.  With >= 2 GPUs each rank owns one and the reduce is NCCL over NVLink (run with
`gpurun --gpus 2`); on a 1-GPU box both ranks share cuda:0 and the reduce goes through gloo (NCCL refuses two ranks on
one device) — the CUDA renderer, the sample split and the sum are exercised either way."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPP, DEPTH, SEED = 24, 16, 3


def _scene(kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from learn_path_tracing_b200 import scenes
    if kind == "v2":
        return scenes.scene_10_final((160, 90))
    from helpers import synthetic_legacy_world
    return synthetic_legacy_world()


def _worker(rank, ws, port, backend, devices, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(devices[rank]))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(devices[rank])
    kw = {"device_id": torch.device("cuda", devices[rank])} if backend == "nccl" else {}
    dist.init_process_group(backend, rank=rank, world_size=ws, **kw)
    import learn_path_tracing_b200 as L
    for kind in ("v2", "legacy"):
        world, cam = _scene(kind)
        for bands in (1, 3):
            img = L.render_distributed(world, cam, spp=SPP, propagate_limit=DEPTH, seed=SEED, postprocess=False, bands=bands)
            assert (img is None) == (rank != 0)
            if rank == 0:
                np.save(os.path.join(outdir, f"{kind}_bands{bands}.npy"), img)
    # progressive legacy passes split over the ranks: 2 x 12 spp == one 24-spp image
    world, cam = _scene("legacy")
    lr = L.legacy.LegacyRenderer(world, cam, spp=SPP // 2, propagate_limit=DEPTH, seed=SEED)
    f1 = lr.render(moved=True)
    f2 = lr.render(moved=False)
    if rank == 0:
        assert lr.total_spp == SPP and f1 is not None
        np.save(os.path.join(outdir, "legacy_progressive.npy"), f2)
    else:
        assert f1 is None and f2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_image_equals_single_gpu_image(ctx, tmp_path):
    import torch
    import torch.multiprocessing as mp
    import learn_path_tracing_b200 as L
    n_dev = torch.cuda.device_count()
    backend, devices = ("nccl", [0, 1]) if n_dev >= 2 else ("gloo", [0, 0])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, backend, devices, str(tmp_path)), nprocs=2, join=True)
    for kind in ("v2", "legacy"):
        world, cam = _scene(kind)
        W, H = cam.resolution
        model = L.PT_SHADE_LEGACY if kind == "legacy" else L.PT_SHADE_V2
        r = L.Renderer(W, H, ctx)
        r.render(world.device_scene(ctx), cam.to_struct(), SPP, DEPTH, model, seed=SEED)
        one = r.mean()
        for bands in (1, 3):
            two = np.load(os.path.join(str(tmp_path), f"{kind}_bands{bands}.npy"))
            assert two.shape == one.shape
            # the same set of paths (RNG keyed on the absolute sample index): equal up to fp32 summation order
            assert np.allclose(two, one, rtol=2e-4, atol=2e-5), (kind, bands, float(np.abs(two - one).max()))
        print(f"{kind}: 2-rank ({backend}) image == 1-GPU image, max |diff| {float(np.abs(two - one).max()):.2e}")
    world, cam = _scene("legacy")
    one = L.legacy.LegacyRenderer(world, cam, spp=SPP, propagate_limit=DEPTH, seed=SEED, ctx=ctx, distributed=False).render()
    two = np.load(os.path.join(str(tmp_path), "legacy_progressive.npy"))
    assert np.allclose(two, one, rtol=2e-3, atol=2e-3, equal_nan=True)
