#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json metric: Mpaths/s, Mrays/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one complete render of the workload (every sample of every pixel) through pt_render.
Default workload = BASELINE.json configs[1]: taichi_pathtracer/8_refract at 1920x1080, 256 spp, depth 50.
N > 1 (torchrun, one rank per GPU): the scene is replicated, every rank renders its own range of sample
indices (256 spp per GPU, weak scaling) and the per-GPU accumulators are summed onto rank 0 by one NCCL
reduce inside the timed region.

--impl reference times the CPU oracle (the reference's algorithm restated in C, all host cores) on a
bounded sample of the same workload; the reference itself (Python + Taichi) cannot be installed here.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, W, H, spp, depth)
    "8_refract_1080p": ("8_refract", 1920, 1080, 256, 50),      # BASELINE configs[1]: the headline workload
    "10_final_720p": ("10_final", 1280, 720, 256, 32),          # configs[0] scene at reduced spp (8192 in the script)
    "10_final_720p_8192": ("10_final", 1280, 720, 8192, 32),    # configs[0] exactly: the script's default resolution/spp/depth
    "9_dof_720p": ("9_dof", 1280, 720, 256, 32),
    # legacy mesh scenes (need scenes_cache/*.npz from tools/prepare_assets.py): configs[2] and configs[3]
    "yoimiya_1080p": ("cache:yoimiya_ground_full", 1920, 1080, 512, 32),
    "zhongli_4k": ("cache:zhongli_full", 3840, 2160, 64, 32),   # configs[3] at reduced spp (4096 in the config)
    # configs[3] exactly: 4K, 4096 spp IN TOTAL, sample ranges split over the N GPUs (strong scaling)
    "zhongli_4k_4096": ("cache:zhongli_full", 3840, 2160, 4096, 32),
    "ganyu_4k_4096": ("cache:ganyu_full", 3840, 2160, 4096, 32),
}
STRONG = {"zhongli_4k_4096", "ganyu_4k_4096"}   # spp is the job total: each of N ranks renders spp / N
# BASELINE configs[4]: synthetic 10M-triangle scene, 64Mi-ray intersection-only batch
INTERSECT = {"intersect_10m": (10_000_000, 64 * 2**20, 12345, 54321, 0.004),
             "intersect_1m": (1_000_000, 8 * 2**20, 12345, 54321, 0.0086)}
# algorithmic HBM bytes: per ray segment / per path
B_EXTEND_SEG = 48    # read o|d 32 B, write hit 16 B
B_SHADE_SEG = 112    # read o|d 32 + throughput 16 + hit 16, write compacted successor 48
B_PER_PATH = 24      # fp32 RGB accumulate read-modify-write
L2_BYTES = 126e6


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with NVML every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_workload(name):
    """-> (world, camera, W, H, spp, depth, shading model, oracle scene builder)"""
    import learn_path_tracing_b200 as L
    from learn_path_tracing_b200 import scenes
    scene, W, H, spp, depth = WORKLOADS[name]
    if scene.startswith("cache:"):
        from learn_path_tracing_b200 import legacy, scene_cache
        path = os.path.join(ROOT, "scenes_cache", scene[6:] + ".npz")
        if not os.path.exists(path):
            raise SystemExit(f"bench.py: {path} missing (python tools/prepare_assets.py needs the reference checkout)")
        world = scene_cache.load_cache(path)
        cam = legacy.Camera((W, H))       # 15_module.py:1068-1072
        cam.set_fov(30)
        cam.set_position(legacy.Vec3f([0, 8, -30]))
        cam.look_at(legacy.Vec3f([0, 8, 0]))
        return world, cam, W, H, spp, depth, L.PT_SHADE_LEGACY
    world, cam = scenes.SCENES[scene]((W, H))
    return world, cam, W, H, spp, depth, L.PT_SHADE_V2


def oracle_scene(world, model):
    from oracle import ptoracle as O
    import learn_path_tracing_b200 as L
    return O.scene_from_legacy_world(world) if model == L.PT_SHADE_LEGACY else O.scene_from_world(world)


def cpu_baseline_run(world, cam, W, H, depth, target_seconds, threads=0, model=0):
    """Times the oracle (kind 'port': the reference's algorithm in C + OpenMP) on a bounded sample."""
    from oracle import ptoracle as O
    sc = oracle_scene(world, model)
    cs = cam.to_struct()
    t0 = time.perf_counter()
    O.render(sc, cs, W, H, 1, depth, model, seed=1, threads=threads)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(1024, target_seconds / max(t1, 1e-3))))
    t0 = time.perf_counter()
    _, _, st = O.render(sc, cs, W, H, spp, depth, model, seed=1, threads=threads)
    dt = time.perf_counter() - t0
    cores = O.num_threads() if threads <= 0 else threads
    return {"mpaths": st.paths / dt / 1e6, "mrays": st.segments / dt / 1e6, "seconds": dt, "spp": spp, "cores": cores}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world, cam, W, H, spp, depth, model = build_workload(args.workload)
    per_step = 8.0
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline_run(world, cam, W, H, depth, per_step if i else 2.0, model=model)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["mpaths"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals]) * 1e3)
    sample = f"{args.workload}: {W}x{H}, {vals[-1]['spp']} spp per step (of {spp}), depth {depth}"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": WORKLOADS[args.workload][0], "width": W, "height": H,
                   "spp": spp, "max_depth": depth,
                   "note": "CPU restatement of the reference algorithm (Taichi is not installable); rate on a bounded sample"},
        "mrays_per_s": float(np.mean([r["mrays"] for r in vals])),
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": vals[-1]["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import learn_path_tracing_b200 as L

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the path tracer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world, cam, W, H, spp, depth, model = build_workload(args.workload)
    strong = args.workload in STRONG
    if strong:
        assert spp % world_size == 0
        spp //= world_size
    ctx = L.default_context()
    scene = world.device_scene(ctx)
    cs = cam.to_struct()
    r = L.Renderer(W, H, ctx)
    flush = torch.empty(int(2 * L2_BYTES) // 4, dtype=torch.float32, device="cuda")
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)

    def step(flags=0):
        flags |= L.PT_FLAG_WIDE if args.wide else 0
        r.clear()
        st = r.render(scene, cs, spp, depth, model, seed=1, spp_offset=rank * spp, flags=flags, mode=args.mode,
                      pool_capacity=args.pool, segments_per_launch=args.k, shade_min=args.shade_min, serve_min=args.serve_min)
        if world_size > 1:
            dist.reduce(r.accum, dst=0, op=dist.ReduceOp.SUM)
        return st

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    times, stats = [], []
    for _ in range(args.steps):
        flush.fill_(1.0)  # evict L2 between timed iterations (the 470 MB path pool exceeds L2 anyway)
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        ev0.record()
        st = step(L.PT_FLAG_TIMING)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
        stats.append(st)
    clocks = sampler.result()
    t_local = float(sum(times))
    if world_size > 1:
        tt = torch.tensor([t_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
        seg = torch.tensor([float(sum(s.segments for s in stats))], dtype=torch.float64, device="cuda")
        dist.all_reduce(seg, op=dist.ReduceOp.SUM)
        seg_total = float(seg.item())
    else:
        t_max, seg_total = t_local, float(sum(s.segments for s in stats))
    paths_total = float(W) * H * spp * args.steps * world_size

    # ---- end-to-end through the public API with host buffers (scene upload + build + render + D2H image)
    e2e = None
    if True:
        legacy_scene = model == L.PT_SHADE_LEGACY
        if legacy_scene:
            h2d = sum(m["positions"].nbytes + m["normals"].nbytes + m["texture_coords"].nbytes + m["indices"].nbytes
                      for m in world.meshes) + world._atlas[0].nbytes + (world._env[0].nbytes if world._env else 0) + 128
        else:
            cr, mats = world.arrays()
            h2d = cr.nbytes + mats.nbytes + 64 + 64
        d2h = W * H * 3 * 4
        e_times = []
        for i in range(2 + min(max(args.steps, 3), 5)):
            world._scene = None  # force re-upload + rebuild: the scene starts on the host every step
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if legacy_scene:
                lr = L.legacy.LegacyRenderer(world, cam, spp=spp, propagate_limit=depth, ctx=ctx)
                img = lr.render(moved=True)
            elif world_size > 1:
                img = L.render_distributed(world, cam, spp=spp * world_size, propagate_limit=depth, seed=1, ctx=ctx)
            else:
                img = L.render(world, cam, spp=spp, propagate_limit=depth, seed=1, ctx=ctx)
            dt = time.perf_counter() - t0
            if i >= 2:  # two untimed calls: pinned-host and device caching allocators warm up
                e_times.append(dt)
            print(f"[bench] e2e iter {i}: {dt*1e3:.2f} ms", file=sys.stderr)
        if rank == 0:
            assert img.shape == (W, H, 3) and np.isfinite(img).all()
        e_t = float(np.median(e_times))   # 3-5 timed calls; the median shrugs off a host hiccup (allocator, nvidia-smi sampler)
        if world_size > 1:
            et = torch.tensor([e_t], dtype=torch.float64, device="cuda")
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            e_t = float(et.item())
        e2e = {"value": W * H * spp * world_size / e_t / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h)}

    if rank == 0:
        peak, peak_kind = measured_peaks()
        K = args.steps
        seg_rank0 = float(sum(s.segments for s in stats))
        paths_rank0 = float(W) * H * spp * K
        ms_ext = float(sum(s.ms_extend for s in stats))
        ms_sh = float(sum(s.ms_shade for s in stats))
        n_ext = int(sum(s.launches_extend for s in stats))
        n_sh = int(sum(s.launches_shade for s in stats))
        if n_ext == 0:  # fused wavefront / persistent kernel: one kernel runs extend + shade + regeneration
            kname = {3: "k_paths_persist", 4: "k_paths_queue", 5: "k_paths_dual"}.get(int(stats[0].reserved[0]), "k_paths")
            bytes_k, ms_k, n_k = 160.0 * seg_rank0 + B_PER_PATH * paths_rank0, ms_sh, n_sh
        elif ms_sh >= ms_ext:
            kname, bytes_k, ms_k, n_k = "k_shade", B_SHADE_SEG * seg_rank0 + B_PER_PATH * paths_rank0, ms_sh, n_sh
        else:
            kname, bytes_k, ms_k, n_k = "k_extend", B_EXTEND_SEG * seg_rank0, ms_ext, n_ext
        achieved = bytes_k / (ms_k * 1e-3) / 1e9
        whole = (160.0 * seg_rank0 + B_PER_PATH * paths_rank0) / (t_local * 1e-3) / 1e9
        try:
            fp32_peak = ctx.measure_fp32_peak()
        except Exception:
            fp32_peak = None
        # one extra untimed render with the counting kernel variant: BVH nodes visited / primitives tested per segment
        r.clear()
        stc = r.render(scene, cs, spp, depth, model, seed=1, spp_offset=rank * spp, flags=L.PT_FLAG_COUNTERS, mode=args.mode,
                       pool_capacity=args.pool, segments_per_launch=args.k, shade_min=args.shade_min, serve_min=args.serve_min)
        nodes_seg = stc.nodes_visited / max(stc.segments, 1)
        prims_seg = stc.prims_tested / max(stc.segments, 1)
        # SURVEY 8d FP32 accounting: two child-box slab tests per BVH2 node (18 flop each), 17 flop per sphere test
        # (45 per Moller-Trumbore triangle test), ~280 flop of shading per segment (transcendentals counted as 1)
        flop_prim = 45.0 if model == L.PT_SHADE_LEGACY else 17.0
        flops_seg = nodes_seg * 36.0 + prims_seg * flop_prim + 280.0
        fp32_achieved = seg_rank0 * flops_seg / (t_local * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"{args.workload}:{kname}")
        # CPU baseline: the oracle on the host cores, a bounded sample; on rank 0 at N = 1 only
        cpu = cpu_baseline_run(world, cam, W, H, depth, 0.5 if args.no_cpu else 12.0, model=model) if world_size == 1 else None
        line = {
            "metric": "Mpaths/s", "value": paths_total / (t_max * 1e-3) / 1e6, "unit": "Mpaths/s",
            "n_gpus": world_size, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": t_max / K,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "scene": WORKLOADS[args.workload][0], "width": W, "height": H,
                       "spp_per_gpu": spp, "max_depth": depth, "parallelism": f"sample-split x{world_size} + NCCL reduce",
                       "l2": "flushed between timed steps (252 MB fill)", "mode": int(stats[0].reserved[0])},
            "mrays_per_s": seg_total / (t_max * 1e-3) / 1e6,
            "segments_per_path": seg_rank0 / paths_rank0,
            "e2e": e2e,
            "gpu_launches": int(sum(s.launches for s in stats)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None, "peak_kind": peak_kind,
                         "launches": n_k, "avg_launch_ms": ms_k / max(n_k, 1),
                         "algorithmic_bytes_per_launch": bytes_k / max(n_k, 1),
                         "algorithmic_bytes": "160 B/segment + 24 B/path for the whole wavefront step (k_paths_persist, k_paths); split mode: k_shade 112 B/segment + 24 B/path, k_extend 48 B/segment (SURVEY 8d)",
                         "whole_render_160B_per_segment": {"achieved": whole, "frac": whole / peak},
                         "kernel_ms": ({kname: ms_sh / K, "step": t_local / K} if n_ext == 0 else
                                       {"k_extend": ms_ext / K, "k_shade": ms_sh / K, "step": t_local / K})},
            "fp32_peak_tflops_measured": fp32_peak,
            "roofline_fp32": {"bound": "fp32", "achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                              "frac": (fp32_achieved / fp32_peak) if fp32_peak else None,
                              "flops_per_segment": flops_seg, "nodes_per_segment": nodes_seg, "prims_per_segment": prims_seg,
                              "accounting": "36 flop per BVH2 node visit + 17 per sphere / 45 per triangle test + 280 shading (SURVEY 8d)"},
            "cpu_baseline": None if cpu is None else {
                "value": cpu["mpaths"], "unit": "Mpaths/s", "cores": cpu["cores"], "kind": "port",
                "sample": f"{W}x{H}, {cpu['spp']} spp, depth {depth}, {cpu['seconds']:.1f} s of OpenMP C oracle (reference algorithm: "
                          + ("unpruned stack walk of the stored SAH tree, texture fetch per candidate"
                             if model == L.PT_SHADE_LEGACY else "brute-force sphere loop per bounce") + ")",
                "mrays_per_s": cpu["mrays"]},
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_intersect(args):
    """BASELINE configs[4]: fixed ray batch against an LBVH over random triangles, intersection only.
    Rays and triangles are generated on the device by counter-based generators (bit-identical to the oracle's);
    N > 1 shards the ray batch by ranges, no collective."""
    import torch
    import torch.distributed as dist
    import learn_path_tracing_b200 as L
    from oracle import ptoracle as O

    n_tri, n_rays, seed_t, seed_r, edge = INTERSECT[args.workload]
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = L.default_context()
    sc = L.Scene(ctx)
    t0 = time.perf_counter()
    sc.set_random_triangles(n_tri, seed_t, edge)
    sc.build()
    build_s = time.perf_counter() - t0
    n_local = n_rays // world_size
    rays = torch.empty((2 * n_rays, 4), dtype=torch.float32, device="cuda")
    ctx.random_rays_device(rays.data_ptr(), n_rays, seed_r)
    my = rays[2 * rank * n_local: 2 * (rank + 1) * n_local]
    hits = torch.empty((n_local, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(int(2 * L2_BYTES) // 4, dtype=torch.float32, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(args.warmup, 3)):
        ctx.trace_batch_device(sc, my.data_ptr(), n_local, hits.data_ptr(), args.trace_flags)
    sampler = ClockSampler(local_rank)
    sampler.start()
    times = []
    for _ in range(args.steps):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        ev0.record()
        ctx.trace_batch_device(sc, my.data_ptr(), n_local, hits.data_ptr(), args.trace_flags)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
    clocks = sampler.result()
    t_local = float(sum(times))
    t_max = t_local
    if world_size > 1:
        tt = torch.tensor([t_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
    st = ctx.trace_batch_device(sc, my.data_ptr(), n_local, hits.data_ptr(), L.PT_FLAG_COUNTERS | args.trace_flags)
    st_plain = ctx.trace_batch_device(sc, my.data_ptr(), n_local, hits.data_ptr(), args.trace_flags)  # sort / traversal split
    # e2e: host ray buffer in, host ids/t out (pt_trace_batch), on a bounded slice
    n_e = min(n_local, 16 * 2**20)
    rays_h = my[:2 * n_e].cpu().numpy().reshape(n_e, 8)
    ids_h, t_h = np.zeros(n_e, np.int32), np.zeros(n_e, np.float32)   # the caller's result arrays, reused every call
    te = []
    for i in range(4):
        t0 = time.perf_counter()
        ctx.trace_batch(sc, rays_h, out=(ids_h, t_h))  # no counters: the plain kernel variant, no per-chunk sync
        if i:  # the first call allocates the pinned / device staging of the context
            te.append(time.perf_counter() - t0)
    if rank == 0:
        peak, peak_kind = measured_peaks()
        K = args.steps
        nodes_per_ray = st.nodes_visited / n_local
        tris_per_ray = st.prims_tested / n_local
        node_bytes = 128 if (args.trace_flags & L.PT_FLAG_TRACE_WIDE) else 64   # 4-wide nodes are 128 bytes
        bytes_per_ray = 32 + 8 + node_bytes * nodes_per_ray + 48 * tris_per_ray
        achieved = bytes_per_ray * n_local * K / (t_local * 1e-3) / 1e9
        kname = "k_trace" if (args.trace_flags & L.PT_FLAG_TRACE_SIMPLE) else "k_trace_persist"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"{args.workload}:{kname}")
        # CPU baseline: the oracle walking the SAME LBVH with the reference triangle test, bounded ray sample
        n_c = 2**18
        nodes, _ = sc.bvh_download()
        tris = O.random_triangles(n_tri, seed_t, edge)
        t0 = time.perf_counter()
        oid, ot, _ = O.trace_bvh2(nodes, tris, rays_h[:n_c])
        cpu_s = time.perf_counter() - t0
        gid = ids_h[:n_c]
        agree = float((gid == oid).mean())
        line = {
            "metric": "Mrays/s", "value": n_local * world_size * K / (t_max * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world_size, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": t_max / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "triangles": n_tri, "rays": n_rays, "edge_scale": edge,
                       "parallelism": f"ray ranges x{world_size}, no collective", "l2": "flushed between timed steps",
                       "lbvh_build_s": build_s},
            "hit_fraction": float((ids_h >= 0).mean()), "ids_equal_to_oracle": agree,
            "e2e": {"value": n_e / min(te) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(n_e * 32),
                    "d2h_bytes_per_step": int(n_e * 8)},   # ids (i32) + t (f32); records are unpacked on the device
            "gpu_launches": K, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None, "peak_kind": peak_kind,
                         "step_ms": {"ray_sort": st_plain.ms_other, "traversal": st_plain.ms_extend, "step": t_local / K},
                         "algorithmic_bytes_per_launch": bytes_per_ray * n_local,
                         "algorithmic_bytes": f"32 + 8 + {node_bytes}*{nodes_per_ray:.1f} nodes + 48*{tris_per_ray:.2f} triangles "
                                              f"= {bytes_per_ray:.0f} B/ray (SURVEY 8d; counts from a counter-instrumented run)",
                         "compulsory_40B_per_ray": {"achieved": 40.0 * n_local * K / (t_local * 1e-3) / 1e9}},
            "cpu_baseline": {"value": n_c / cpu_s / 1e6, "unit": "Mrays/s", "cores": O.num_threads(), "kind": "port",
                             "sample": f"{n_c} rays of the same batch, oracle walking the same LBVH with the reference "
                                       f"triangle test, {cpu_s:.1f} s"},
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="8_refract_1080p", choices=sorted(WORKLOADS) + sorted(INTERSECT))
    ap.add_argument("--mode", type=int, default=0, help="wavefront mode: 0 auto (= 3 persistent), 1 split kernels, 2 K-step fused, 3 persistent, 4 queue, 5 dual")
    ap.add_argument("--pool", type=int, default=0, help="path-pool slots (0 = library default)")
    ap.add_argument("--k", type=int, default=0, help="fused mode: segments per launch (0 = default); dual mode: blocks per SM (3 or 4)")
    ap.add_argument("--shade-min", type=int, default=0, help="persistent mode: waiting lanes that trigger shading (0 = default)")
    ap.add_argument("--serve-min", type=int, default=0, help="persistent mode: waiting lanes that trigger a service (0 = default)")
    ap.add_argument("--trace-flags", type=int, default=0, help="intersect workloads: PT_FLAG_* for pt_trace_batch_device")
    ap.add_argument("--wide", action="store_true",
                    help="EXPERIMENTAL: build the 4-wide copy of the tree (PT_WIDE=1) and walk it (PT_FLAG_WIDE / PT_FLAG_TRACE_WIDE)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.wide:
        os.environ["PT_WIDE"] = "1"          # read by pt_scene_build
        args.trace_flags |= L.PT_FLAG_TRACE_WIDE
    if args.workload in INTERSECT:
        if args.impl == "reference":
            raise SystemExit("--impl reference: use a render workload (the intersect line carries its own cpu_baseline)")
        run_intersect(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
