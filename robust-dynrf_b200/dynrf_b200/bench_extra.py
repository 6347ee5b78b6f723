"""bench.py legs for BASELINE configs 4 and 5 (extra evidence lines; the default bench line stays cfg 2):

  run_train9  --config 4 --mode train9 : the reference's full 9-pass train step (dynrf_b200/train9.py), rays/s, weak or
                                         strong (--strong: 4096 GLOBAL rays / N) scaling
  run_render  --config 5               : frame-sharded 1080p render path (dynrf_b200/render_path.py), frames/s
Same timing rules as bench.py: >= 3 warm-up steps, CUDA events, L2 flush between timed iterations, max over ranks, clocks."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch


def _barrier(dist):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def run_train9(args, w, rank, world, dev, dist, ClockSampler, config_dict, METRIC):
    from . import synth as SY
    from . import train9 as T9

    strong = bool(int(os.environ.get("DYNRF_STRONG", "0")))
    n_rays = w["n_rays"] // world if strong else w["n_rays"]
    Himg, Wimg, F = 270, 480, 24
    static, dyn = SY.make_models(w["grid"], w["ray_type"], w["static_mode"], dev, seed=1)
    step = T9.Train9(static, dyn, w["n_samples"], w["ray_type"], Himg, Wimg, F, focal=418.0 / 2, world=world, graph=not args.no_graph)
    g = torch.Generator().manual_seed(2000 + rank)
    batch = T9.synth_batch(n_rays, Himg, Wimg, F, g, dev)
    host = {k: v.cpu().pin_memory() for k, v in batch.items()}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out_pin = torch.empty(n_rays, 5, pin_memory=True)
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step(batch)
    _barrier(dist)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        step(batch)
        ev[i][1].record()
    _barrier(dist)
    it = [a.elapsed_time(b) for a, b in ev]
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev2[i][0].record()
        b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        o = step(b)
        out_pin.copy_(o, non_blocking=True)
        ev2[i][1].record()
    _barrier(dist)
    it2 = [a.elapsed_time(b) for a, b in ev2]
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([sum(it), sum(it2)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        t_ms, t2_ms = float(tt[0]), float(tt[1])
        total = n_rays * world * args.steps
        cfg = config_dict(w, "train9", world, not args.no_graph)
        cfg.update(rays_per_gpu=n_rays, global_rays=n_rays * world,
                   passes="9 hot-path passes of train.py:1092-2292 (dynrf_b200/train9.py); flatten_eff_distloss omitted (third party, absent)",
                   image="synthetic 24 frames of 480x270, focal 209; joint pose (9-D) + focal optimisation")
        line = {"metric": METRIC + " [full 9-pass step]", "value": total / (t_ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "e2e": {"value": total / (t2_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
                        "d2h_bytes_per_step": int(out_pin.numel() * 4), "ms_per_step": t2_ms / args.steps},
                "ms_per_iter": {"value": [round(x, 3) for x in it], "e2e": [round(x, 3) for x in it2]},
                "gpu_launches": (step.launches_per_step or 0) * args.steps, "clocks": clocks}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_render(args, w, rank, world, dev, dist, ClockSampler, config_dict):
    from . import synth as SY
    from . import render_path as RP

    Himg, Wimg, n_frames, focal = 1080, 1920, 270, 1000.0
    static, dyn = SY.make_models(w["grid"], "ndc", w["static_mode"], dev, seed=1)
    drv = RP.RenderDriver(static, dyn, Himg, Wimg, focal, w["n_samples"], "ndc", chunk=8192, rank=rank, world=world, graph=not args.no_graph)
    poses = RP.change_view_time_poses(np.concatenate([np.eye(3), np.zeros((3, 1))], 1), max_trans=0.05, length=n_frames)
    mine = drv.my_frames(n_frames)
    k = args.frames if args.frames > 0 else min(len(mine), max(1, args.steps // 10))
    warm, timed = mine[:1], mine[1:1 + k]
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    drv.render(poses, frames=warm)       # graph capture + warm-up (254 chunk replays)
    _barrier(dist)
    drv.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = drv.render(poses, frames=timed)
    e1.record()
    _barrier(dist)
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms, wall_ms = float(tt[0]), float(tt[1])
        frames_total = len(timed) * world
        cfg = config_dict(w, "render", world, not args.no_graph)
        cfg.update(rays_per_gpu=Himg * Wimg, ray_order="render (row-major pixels of a frame, renderer.py:1055)", frames_timed_per_gpu=len(timed),
                   frames_total=n_frames, image=f"{Wimg}x{Himg}", chunk=8192, parallelism=f"frames round-robin over {world} rank(s), no collective",
                   l2="no flush: a frame streams 2.07 M rays / 531 M samples through the kernels (working set >> L2)")
        line = {"metric": "frames/sec (render_path 1080p, change_view_time)", "value": frames_total / (wall_ms * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": len(timed), "warmup": len(warm), "ms_per_step": wall_ms / max(len(timed), 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "rays_per_s": frames_total * Himg * Wimg / (wall_ms * 1e-3),
                "device_ms_per_frame": ms / max(len(timed), 1),
                "e2e": {"value": frames_total / (wall_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 48 + 4,
                        "d2h_bytes_per_step": Himg * Wimg * 16 * 4, "note": "poses in, all 16 per-ray maps of every frame back in pinned host memory (async, double buffered); wall clock"},
                "gpu_launches": drv.launches, "clocks": clocks}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
